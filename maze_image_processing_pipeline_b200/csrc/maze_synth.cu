// Synthetic LOKI-shaped vignettes generated on the device (benchmark input only).
// Same image model as maze_image_processing_pipeline_b200/synth.py (dark N(12,4) background plus
// 1-6 rotated anisotropic Gaussian blobs) with a counter-based hash RNG keyed by
// (seed, global vignette index, pixel), so any slice of the job can be regenerated anywhere.
#include <math.h>

#include "maze_common.cuh"

__device__ __forceinline__ u64 mix64(u64 z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(u64 h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(MAZE_CTA) k_synth(uint8_t *__restrict__ image, const maze_vignette_t *__restrict__ vig,
                                                    const maze_tile_t *__restrict__ tiles, u64 seed, i64 img_index0)
{
    __shared__ float s_blob[6][6];
    TileCtx c = load_tile(vig, tiles);
    const int H = c.v.h, W = c.v.w;
    u64 key = mix64(seed ^ mix64((u64)(img_index0 + c.img)));
    if (threadIdx.x < 6) {
        int b = threadIdx.x;
        u64 kb = mix64(key + 1000 + b);
        int nb = 1 + (int)(mix64(key + 7) % 6);
        float cy = (0.15f + 0.7f * u01(mix64(kb + 1))) * H, cx = (0.15f + 0.7f * u01(mix64(kb + 2))) * W;
        float sy = (0.02f + 0.08f * u01(mix64(kb + 3))) * H, sx = (0.02f + 0.08f * u01(mix64(kb + 4))) * W;
        float th = 3.14159265f * u01(mix64(kb + 5));
        float amp = b < nb ? 80.0f + 140.0f * u01(mix64(kb + 6)) : 0.0f;
        float ct = cosf(th), st = sinf(th);
        s_blob[b][0] = cy; s_blob[b][1] = cx;
        s_blob[b][2] = 0.5f * (ct * ct / (sy * sy) + st * st / (sx * sx));
        s_blob[b][3] = 0.5f * ct * st * (1.0f / (sy * sy) - 1.0f / (sx * sx));
        s_blob[b][4] = 0.5f * (st * st / (sy * sy) + ct * ct / (sx * sx));
        s_blob[b][5] = amp;
    }
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *img = image + c.v.pix_off;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        if (x < W) {
            u64 h = mix64(key ^ ((u64)y * 65537ull + (u64)x) * 0x2545f4914f6cdd1dull);
            float u1 = u01(h), u2 = u01(mix64(h));
            float val = 12.0f + 4.0f * sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
#pragma unroll
            for (int b = 0; b < 6; b++) {
                float amp = s_blob[b][5];
                if (amp > 0.0f) {
                    float dy = y - s_blob[b][0], dx = x - s_blob[b][1];
                    float q = s_blob[b][2] * dy * dy + 2.0f * s_blob[b][3] * dy * dx + s_blob[b][4] * dx * dx;
                    if (q < 12.0f) val += amp * __expf(-q);
                }
            }
            val = fminf(fmaxf(rintf(val), 0.0f), 255.0f);
            img[(i64)y * W + x] = (uint8_t)val;
        }
        if (++k == c.v.wpr) { k = 0; y++; }
    }
}

extern "C" int maze_synth_vignettes(uint8_t *image, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                                    int n_tiles, uint64_t seed, int64_t img_index0, void *stream)
{
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    MAZE_KERNEL(KID_SYNTH, (cudaStream_t)stream, k_synth<<<n_tiles, MAZE_CTA, 0, (cudaStream_t)stream>>>(image, vig, tiles, (u64)seed, (i64)img_index0));
    return MAZE_OK;
}
