// Threshold / bit-plane packing, radius-bounded thresholded EDT (isotropic morphology on bit
// planes) and the exact squared EDT.  sm_100a.
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:649, :405         threshold, bool cast
//   maze_ipp/isotropic.py:35-36, 66-67          dist > radius / dist < radius on scipy's EDT
#include "maze_common.cuh"

thread_local char maze_err_buf[256] = "";

void maze_set_err(cudaError_t e, const char *where)
{
    snprintf(maze_err_buf, sizeof(maze_err_buf), "%s: %s", where, cudaGetErrorString(e));
}

extern "C" const char *maze_error_string(void) { return maze_err_buf; }
extern "C" int maze_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// threshold + pack: one thread per 32-pixel word (aligned 32-bit loads + byte-SIMD compare, maze_common.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAZE_CTA) k_threshold_pack(const uint8_t *__restrict__ image,
                                                             const maze_vignette_t *__restrict__ vig,
                                                             const maze_tile_t *__restrict__ tiles, int t_int,
                                                             uint32_t *__restrict__ bits, uint32_t *flags)
{
    __shared__ uint32_t s_fl;
    TileCtx c = load_tile(vig, tiles);
    if (threadIdx.x == 0) s_fl = 0;
    __syncthreads();
    const int widx = c.word0 + threadIdx.x;
    uint32_t fl = 0;
    if (widx < c.nwords) {
        const int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
        const uint32_t word = threshold_word32(image + c.v.pix_off, y, k, c.v.w, t_int);
        bits[c.v.word_off + widx] = word;
        const uint32_t vm = valid_mask(c.v.w, k);
        fl = (word ? 1u : 0u) | ((word ^ vm) ? 2u : 0u);
    }
    fl = __reduce_or_sync(FULL, fl);
    if ((threadIdx.x & 31) == 0 && fl) atomicOr(&s_fl, fl);
    __syncthreads();
    if (threadIdx.x == 0 && s_fl) atomicOr(flags + c.img, s_fl);
}

extern "C" int maze_threshold_pack(const uint8_t *image, const maze_vignette_t *vig, int n_img,
                                   const maze_tile_t *tiles, int n_tiles, int t_int, uint32_t *bits,
                                   uint32_t *flags, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    MAZE_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)n_img, s), "threshold flags");
    MAZE_KERNEL(KID_THRESHOLD_PACK, s, k_threshold_pack<<<n_tiles, MAZE_CTA, 0, s>>>(image, vig, tiles, t_int, bits, flags));
    return MAZE_OK;
}

// same packing for an int32 squared-distance map: bit = (d2 > t) or (d2 <= t)
__global__ void __launch_bounds__(MAZE_CTA) k_compare_pack(const int32_t *__restrict__ d2,
                                                           const maze_vignette_t *__restrict__ vig,
                                                           const maze_tile_t *__restrict__ tiles, int t, int greater,
                                                           uint32_t *__restrict__ bits, uint32_t *flags)
{
    TileCtx c = load_tile(vig, tiles);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *src = d2 + c.v.pix_off;
    uint32_t fl = 0;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        bool p = false;
        if (x < c.v.w) {
            int v = src[(i64)y * c.v.w + x];
            p = greater ? (v > t) : (v <= t);
        }
        uint32_t word = __ballot_sync(FULL, p);
        if (lane == 0) {
            bits[c.v.word_off + widx] = word;
            uint32_t vm = valid_mask(c.v.w, k);
            fl |= (word ? 1u : 0u) | ((word ^ vm) ? 2u : 0u);
        }
        if (++k == c.v.wpr) { k = 0; y++; }
    }
    if (lane == 0 && fl) atomicOr(flags + c.img, fl);
}

extern "C" int maze_compare_pack(const int32_t *d2, const maze_vignette_t *vig, int n_img,
                                 const maze_tile_t *tiles, int n_tiles, int t, int greater, uint32_t *bits,
                                 uint32_t *flags, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    MAZE_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)n_img, s), "compare flags");
    MAZE_KERNEL(KID_COMPARE_PACK, s, k_compare_pack<<<n_tiles, MAZE_CTA, 0, s>>>(d2, vig, tiles, t, greater, bits, flags));
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------
// unpack: bit plane -> bool bytes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAZE_CTA) k_unpack_mask(const uint32_t *__restrict__ bits,
                                                          const maze_vignette_t *__restrict__ vig,
                                                          const maze_tile_t *__restrict__ tiles,
                                                          uint8_t *__restrict__ mask)
{
    TileCtx c = load_tile(vig, tiles);
    uint8_t *dst = mask + c.v.pix_off;
    if ((c.v.w & 3) == 0) {
        // rows start 4-byte aligned: one thread per word, each nibble expanded to four 0/1 bytes by a multiply
        const int widx = c.word0 + threadIdx.x;
        if (widx >= c.nwords) return;
        const int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
        const uint32_t word = __ldg(bits + c.v.word_off + widx);
        uint32_t *d4 = (uint32_t *)(dst + (size_t)y * c.v.w + 32 * k);
        const int n4 = min(8, (c.v.w - 32 * k) >> 2);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < n4) d4[i] = (((word >> (4 * i)) & 0xfu) * 0x00204081u) & 0x01010101u;
        return;
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        uint32_t word = __ldg(bits + c.v.word_off + widx);
        int x = 32 * k + lane;
        if (x < c.v.w) dst[(i64)y * c.v.w + x] = (word >> lane) & 1u;
        if (++k == c.v.wpr) { k = 0; y++; }
    }
}

extern "C" int maze_unpack_mask(const uint32_t *bits, const maze_vignette_t *vig, int n_img,
                                const maze_tile_t *tiles, int n_tiles, uint8_t *mask, void *stream)
{
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    MAZE_KERNEL(KID_UNPACK_MASK, (cudaStream_t)stream, k_unpack_mask<<<n_tiles, MAZE_CTA, 0, (cudaStream_t)stream>>>(bits, vig, tiles, mask));
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------
// Radius-bounded thresholded EDT on bit planes.
//
// isotropic_erosion keeps pixel p iff d2(p) > t, i.e. iff no background pixel lies at an offset
// (dy, dx) with dy^2 + dx^2 <= t.  Separably: for every |dy| <= R = isqrt(t) the row y+dy must be
// foreground on the whole horizontal chord |dx| <= w(dy) = isqrt(t - dy^2).  On bit planes a chord
// test is an AND of funnel-shifted words and the column combination is an AND over dy -- the EDT's
// row and column passes evaluated only inside the radius and never materialising d2.
// Dilation is the same test on the complement plane (out-of-image pixels count as "no
// foreground" there and as "no background" for erosion, so both read 1 after inversion).
// A plane with no 0 after inversion gets scipy's phantom background pixel at (-1, 0).
// ---------------------------------------------------------------------------------------------
struct DiskTab {
    int R;                              // isqrt(t), or -1 for the empty disk
    int w[MAZE_MAX_DISK_RADIUS + 1];    // half chord per |dy|
    int use_phantom;                    // EDT semantics (scipy's phantom pixel) or plain binary morphology
};

// ---- footprint registry: symmetric, row-convex footprints given as half chords per |dy| (skimage.morphology.disk
// with or without decomposition="crosses" collapses to one such footprint, loki/pipeline.py:408-427) ------------
#include <mutex>
#define MAZE_MAX_FOOTPRINTS 256
static std::mutex g_fp_mutex;
static int g_fp_count = 0;
static int g_fp_R[MAZE_MAX_FOOTPRINTS];
static int g_fp_w[MAZE_MAX_FOOTPRINTS][MAZE_MAX_DISK_RADIUS + 1];

extern "C" int maze_footprint_register(int R, const int32_t *w)
{
    if (R < 0 || R > MAZE_MAX_DISK_RADIUS || !w) return MAZE_ERR_BADARG;
    for (int i = 0; i <= R; i++)
        if (w[i] < 0 || w[i] > MAZE_MAX_DISK_RADIUS) return MAZE_ERR_BADARG;
    std::lock_guard<std::mutex> lock(g_fp_mutex);
    for (int id = 0; id < g_fp_count; id++) {
        if (g_fp_R[id] != R) continue;
        bool same = true;
        for (int i = 0; i <= R && same; i++) same = g_fp_w[id][i] == w[i];
        if (same) return id;
    }
    if (g_fp_count >= MAZE_MAX_FOOTPRINTS) return MAZE_ERR_CAPACITY;
    const int id = g_fp_count;
    g_fp_R[id] = R;
    for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) g_fp_w[id][i] = i <= R ? w[i] : 0;
    g_fp_count = id + 1;  // published last: readers only look below the count
    return id;
}

static int isqrt_tab(int v)
{
    int r = 0;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}

bool maze_pass_table(int t, int *R, int *w, int *use_phantom)
{
    for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) w[i] = 0;
    *use_phantom = 1;
    if (t == -1) { *R = -1; return true; }
    if (t >= 0) {
        if (t >= (MAZE_MAX_DISK_RADIUS + 1) * (MAZE_MAX_DISK_RADIUS + 1)) return false;
        *R = isqrt_tab(t);
        for (int dy = 0; dy <= *R; dy++) w[dy] = isqrt_tab(t - dy * dy);
        return true;
    }
    const int id = -2 - t;  // MAZE_FOOTPRINT_T
    std::lock_guard<std::mutex> lock(g_fp_mutex);
    if (id < 0 || id >= g_fp_count) return false;
    *R = g_fp_R[id];
    for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) w[i] = g_fp_w[id][i];
    *use_phantom = 0;
    return true;
}

__device__ __forceinline__ uint32_t morph_load(const uint32_t *plane, int H, int W, int wpr, int yy, int kk,
                                               uint32_t inv, bool phantom)
{
    if (kk < 0 || kk >= wpr) return FULL;
    if (yy < 0 || yy >= H) return (phantom && yy == -1 && kk == 0) ? 0xfffffffeu : FULL;
    uint32_t v = __ldg(plane + (i64)yy * wpr + kk) ^ inv;
    return v | ~valid_mask(W, kk);
}

__global__ void __launch_bounds__(MAZE_CTA) k_morph_pass(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                         const maze_vignette_t *__restrict__ vig,
                                                         const maze_tile_t *__restrict__ tiles, DiskTab disk,
                                                         int invert, const uint32_t *__restrict__ flags_in,
                                                         uint32_t *flags_out)
{
    __shared__ uint32_t s_fl;
    TileCtx c = load_tile(vig, tiles);
    if (threadIdx.x == 0) s_fl = 0;
    __syncthreads();
    int widx = c.word0 + threadIdx.x;
    uint32_t fl = 0;
    if (widx < c.nwords) {
        const int H = c.v.h, W = c.v.w, wpr = c.v.wpr;
        const uint32_t *plane = in + c.v.word_off;
        int y = widx / wpr, k = widx - y * wpr;
        uint32_t inv = invert ? FULL : 0u;
        bool phantom = false;
        if (disk.use_phantom) {
            uint32_t fin = flags_in[c.img];
            phantom = invert ? !(fin & 1u) : !(fin & 2u);
        }
        uint32_t acc = FULL;
        for (int dy = -disk.R; dy <= disk.R; dy++) {
            int yy = y + dy;
            int w = disk.w[dy < 0 ? -dy : dy];
            uint32_t C = morph_load(plane, H, W, wpr, yy, k, inv, phantom);
            uint32_t h = C;
            if (w > 0) {
                uint32_t L = morph_load(plane, H, W, wpr, yy, k - 1, inv, phantom);
                uint32_t Rw = morph_load(plane, H, W, wpr, yy, k + 1, inv, phantom);
                for (int d = 1; d <= w; d++) {
                    h &= __funnelshift_rc(C, Rw, d); // bit x <- pixel x + d
                    h &= __funnelshift_lc(L, C, d);  // bit x <- pixel x - d
                }
            }
            acc &= h;
        }
        uint32_t vm = valid_mask(W, k);
        uint32_t res = (acc ^ inv) & vm;
        out[c.v.word_off + widx] = res;
        fl = (res ? 1u : 0u) | ((res ^ vm) ? 2u : 0u);
    }
    fl = __reduce_or_sync(FULL, fl);
    if ((threadIdx.x & 31) == 0 && fl) atomicOr(&s_fl, fl);
    __syncthreads();
    if (threadIdx.x == 0 && s_fl) atomicOr(flags_out + c.img, s_fl);
}

extern "C" int maze_morph_pass(const uint32_t *in, uint32_t *out, const maze_vignette_t *vig, int n_img,
                               const maze_tile_t *tiles, int n_tiles, int t, int invert, const uint32_t *flags_in,
                               uint32_t *flags_out, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    if (in == out) return MAZE_ERR_BADARG;
    DiskTab disk;
    if (!maze_pass_table(t, &disk.R, disk.w, &disk.use_phantom)) return MAZE_ERR_BADARG;
    MAZE_CUDA(cudaMemsetAsync(flags_out, 0, sizeof(uint32_t) * (size_t)n_img, s), "morph flags");
    MAZE_KERNEL(KID_MORPH_PASS, s, k_morph_pass<<<n_tiles, MAZE_CTA, 0, s>>>(in, out, vig, tiles, disk, invert, flags_in, flags_out));
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------
// Exact squared EDT: column pass (vertical distance g), then a row pass that stages the row of g
// in shared memory and takes min_k (k^2 + g(x+k)^2) with the scan pruned at k^2 >= best.
// ---------------------------------------------------------------------------------------------
#define G_INF (1 << 24) /* "no zero in this column"; rows, cols <= 2^15 so g^2 of real values fits int32 */

__global__ void k_plane_has_zero(const uint32_t *__restrict__ bits, const maze_vignette_t *__restrict__ vig,
                                 int invert, uint32_t *has_zero)
{
    maze_vignette_t v = vig[blockIdx.x];
    int nwords = v.h * v.wpr;
    uint32_t inv = invert ? FULL : 0u;
    bool z = false;
    for (int i = threadIdx.x; i < nwords && !z; i += blockDim.x) {
        int k = i % v.wpr;
        uint32_t m = (bits[v.word_off + i] ^ inv) | ~valid_mask(v.w, k);
        z = (m != FULL);
    }
    if (__syncthreads_or(z) && threadIdx.x == 0) has_zero[blockIdx.x] = 1;
}

__global__ void k_edt_cols(const uint32_t *__restrict__ bits, const maze_vignette_t *__restrict__ vig, int invert,
                           const uint32_t *__restrict__ has_zero, int32_t *__restrict__ d2)
{
    maze_vignette_t v = vig[blockIdx.x];
    int x = blockIdx.y * blockDim.x + threadIdx.x;
    if (x >= v.w) return;
    const uint32_t *plane = bits + v.word_off + (x >> 5);
    uint32_t inv = invert ? 1u : 0u;
    int sh = x & 31;
    int32_t *g = d2 + v.pix_off + x;
    bool phantom = !has_zero[blockIdx.x];
    int last = (phantom && x == 0) ? -1 : -G_INF;
    for (int y = 0; y < v.h; y++) {
        uint32_t b = ((plane[(i64)y * v.wpr] >> sh) & 1u) ^ inv;
        if (!b) last = y;
        int d = y - last;
        g[(i64)y * v.w] = d > G_INF ? G_INF : d;
    }
    last = G_INF;
    for (int y = v.h - 1; y >= 0; y--) {
        uint32_t b = ((plane[(i64)y * v.wpr] >> sh) & 1u) ^ inv;
        if (!b) last = y;
        int d = last - y;
        int cur = g[(i64)y * v.w];
        if (d < cur) g[(i64)y * v.w] = d;
    }
}

__global__ void k_edt_rows(const maze_vignette_t *__restrict__ vig, int32_t *__restrict__ d2)
{
    extern __shared__ int32_t s_g[];
    maze_vignette_t v = vig[blockIdx.x];
    for (int y = blockIdx.y; y < v.h; y += gridDim.y) {
        int32_t *row = d2 + v.pix_off + (i64)y * v.w;
        for (int x = threadIdx.x; x < v.w; x += blockDim.x) s_g[x] = row[x];
        __syncthreads();
        for (int x = threadIdx.x; x < v.w; x += blockDim.x) {
            int g0 = s_g[x];
            i64 best = g0 >= G_INF ? ((i64)1 << 60) : (i64)g0 * g0;
            for (i64 k = 1; k * k < best; k++) {
                bool any = false;
                if (x - k >= 0) {
                    any = true;
                    int gv = s_g[x - k];
                    if (gv < G_INF) {
                        i64 cnd = k * k + (i64)gv * gv;
                        if (cnd < best) best = cnd;
                    }
                }
                if (x + k < v.w) {
                    any = true;
                    int gv = s_g[x + k];
                    if (gv < G_INF) {
                        i64 cnd = k * k + (i64)gv * gv;
                        if (cnd < best) best = cnd;
                    }
                }
                if (!any) break;
            }
            row[x] = (int32_t)(best > 0x7fffffff ? 0x7fffffff : best);
        }
        __syncthreads();
    }
}

extern "C" int maze_edt_sq(const uint32_t *bits, const maze_vignette_t *vig, int n_img, int max_h, int max_w,
                           int invert, int32_t *d2, uint32_t *has_zero, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || max_h <= 0 || max_w <= 0) return MAZE_OK;
    if ((size_t)max_w * 4 > 200 * 1024) return MAZE_ERR_BADARG;
    uint32_t *pool = has_zero;
    MAZE_CUDA(cudaMemsetAsync(pool, 0, sizeof(uint32_t) * (size_t)n_img, s), "edt flags");
    MAZE_KERNEL(KID_PLANE_HAS_ZERO, s, k_plane_has_zero<<<n_img, 256, 0, s>>>(bits, vig, invert, pool));
    dim3 gc(n_img, (max_w + 127) / 128);
    MAZE_KERNEL(KID_EDT_COLS, s, k_edt_cols<<<gc, 128, 0, s>>>(bits, vig, invert, pool, d2));
    int gy = max_h;
    if ((i64)gy * n_img > 65535 * 8) gy = (int)((65535 * 8) / n_img);
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    size_t smem = (size_t)max_w * 4;
    if (smem > 48 * 1024)
        MAZE_CUDA(cudaFuncSetAttribute(k_edt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "edt smem");
    MAZE_KERNEL(KID_EDT_ROWS, s, k_edt_rows<<<dim3(n_img, gy), 256, smem, s>>>(vig, d2));
    return MAZE_OK;
}
