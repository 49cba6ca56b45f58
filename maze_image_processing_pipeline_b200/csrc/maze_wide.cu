// Thresholded EDT for LARGE radii (isotropic erosion / dilation, maze_ipp/isotropic.py:35-36, 66-67; footprint
// morphology, maze_ipp/loki/pipeline.py:408-427) as a SEPARABLE pass pair whose cost grows with R, not R^2:
//
//   A k_wide_vdist   g(y, x) = vertical distance of pixel (y, x) to the nearest 0 of its column (0 on a 0 pixel),
//                    capped at R + 1, one byte per pixel: every thread walks a 32-row segment of one pixel column
//                    down and up with a counter (started R + 1 rows outside the segment, so the capped values are
//                    exact); the 32 lanes of a warp are the 32 columns of one bit-plane word.
//   B k_wide_rows    pixel (y, x) survives iff no 0 lies at an offset inside the disk, i.e. iff for every |dx| <= Rw
//                    g(y, x + dx) > h[|dx|] with h[d] = isqrt(t - d^2) (the half height of the disk over column d; for a
//                    registered footprint the largest |dy| whose chord reaches d).  The row of g is staged in shared
//                    memory; a thread tests its pixel outward from dx = 0 and stops at the first failure.
//
// d^2 is never materialised and no float is involved: the compare is the same integer test as in maze_morph_pass
// (d2 > t  <=>  no background pixel with dy^2 + dx^2 <= t).  Pixels outside the image are foreground (the image
// border is not background, as for scipy's EDT) except for scipy's phantom background pixel at (-1, 0) of a plane
// without any 0.  Dilation is the same test on the complement plane.  Radii up to 254 (g fits a byte).
#include "maze_common.cuh"

#define WIDE_MAX_R 254
#define WIDE_SEG 32   /* rows per thread in pass A */
#define WIDE_CHUNK 256 /* pixels per CTA in pass B */

struct WideTab {
    int R;   // rows the structuring element reaches up / down
    int Rw;  // columns it reaches left / right
    int use_phantom;
    uint8_t h[WIDE_MAX_R + 1]; // h[|dx|]: rows it reaches over column dx
};

__device__ __forceinline__ int wide_find(const int64_t *off, int n, int64_t v)
{
    int lo = 0, hi = n; // largest i with off[i] <= v
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// pass A: one warp = one bit-plane word column x one 32-row segment; CTA = 8 warps = 8 consecutive word columns
__global__ void __launch_bounds__(256) k_wide_vdist(const uint32_t *__restrict__ in, const maze_vignette_t *__restrict__ vig,
                                                    int n_img, const int64_t *__restrict__ cta_off, int R, int invert,
                                                    int use_phantom, const uint32_t *__restrict__ flags_in,
                                                    uint8_t *__restrict__ g, uint32_t *__restrict__ clear)
{
    const int img = wide_find(cta_off, n_img, (int64_t)blockIdx.x);
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, wpr = v.wpr;
    const int kgroups = (wpr + 7) >> 3;
    const int local = (int)((int64_t)blockIdx.x - cta_off[img]);
    const int seg = local / kgroups, k = (local - seg * kgroups) * 8 + (threadIdx.x >> 5);
    if (k >= wpr) return;
    const int lane = threadIdx.x & 31, x = 32 * k + lane;
    const uint32_t inv = invert ? FULL : 0u;
    bool phantom = false;
    if (use_phantom) {
        const uint32_t fin = flags_in[img];
        phantom = invert ? !(fin & 1u) : !(fin & 2u); // the (inverted) input plane has no 0
    }
    const uint32_t *col = in + v.word_off + k;
    const uint32_t pad = ~valid_mask(W, k);
    const int cap = R + 1;
    const int y0 = seg * WIDE_SEG, y1 = min(H, y0 + WIDE_SEG);
    auto bit = [&](int y) { return (((col[(size_t)y * wpr] ^ inv) | pad) >> lane) & 1u; };
    // walk down: distance to the nearest 0 above (or at) the pixel
    int d = cap; // no 0 in reach above the start row
    int ys = y0 - cap;
    if (ys < 0) {
        ys = 0;
        if (phantom && x == 0) d = 0; // scipy's phantom pixel sits right above (0, 0): row -1 is a 0
    }
#pragma unroll 8
    for (int y = ys; y < y0; y++) d = bit(y) ? min(d + 1, cap) : 0;
    int dn[WIDE_SEG];
#pragma unroll
    for (int i = 0; i < WIDE_SEG; i++) {
        const int y = y0 + i;
        if (y < y1) d = bit(y) ? min(d + 1, cap) : 0;
        dn[i] = d;
    }
    // walk up: distance to the nearest 0 below, combined with the one above
    d = cap;
#pragma unroll 8
    for (int y = min(H, y1 + cap) - 1; y >= y1; y--) d = bit(y) ? min(d + 1, cap) : 0;
    uint8_t *go = g + v.pix_off;
    uint32_t *co = clear + v.word_off + k;
#pragma unroll
    for (int i = WIDE_SEG - 1; i >= 0; i--) {
        const int y = y0 + i;
        if (y < y1) {
            d = bit(y) ? min(d + 1, cap) : 0;
            const int gv = min(d, dn[i]);
            if (x < W) go[(size_t)y * W + x] = (uint8_t)gv;
            // "clear" plane: the column holds no 0 within R rows of the pixel (pad columns count as clear)
            const uint32_t cw = __ballot_sync(FULL, gv == cap);
            if (lane == 0) co[(size_t)y * wpr] = cw;
        }
    }
}

// pass B: one CTA = 256 consecutive pixel columns x WIDE_ROWS consecutive rows.  A pixel whose own column is not
// clear fails at once; one whose whole window [x - Rw, x + Rw] is clear (a few word tests on the "clear" plane)
// survives at once; only the pixels in between -- those within reach of a 0 -- run the test against the half heights.
#define WIDE_ROWS 8
#define WIDE_CW (WIDE_CHUNK / 32 + 2 * ((WIDE_MAX_R + 31) / 32) + 2)
__global__ void __launch_bounds__(WIDE_CHUNK) k_wide_rows(const uint8_t *__restrict__ g, const uint32_t *__restrict__ clear,
                                                           const maze_vignette_t *__restrict__ vig, int n_img,
                                                           const int64_t *__restrict__ cta_off, WideTab tab, int invert,
                                                           uint32_t *__restrict__ out, uint32_t *flags_out)
{
    __shared__ uint8_t s_g[WIDE_ROWS][WIDE_CHUNK + 2 * WIDE_MAX_R + 8];
    __shared__ uint32_t s_c[WIDE_ROWS][WIDE_CW];
    __shared__ uint8_t s_h[WIDE_MAX_R + 1];
    __shared__ uint32_t s_fl;
    const int img = wide_find(cta_off, n_img, (int64_t)blockIdx.x);
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, wpr = v.wpr;
    const int chunks = (W + WIDE_CHUNK - 1) / WIDE_CHUNK;
    const int local = (int)((int64_t)blockIdx.x - cta_off[img]);
    const int ys = local / chunks, x0 = (local - ys * chunks) * WIDE_CHUNK;
    const int y0 = ys * WIDE_ROWS, y1 = min(H, y0 + WIDE_ROWS);
    const int Rw = tab.Rw;
    const int gw = (Rw + 31) >> 5;              // guard words on each side of the chunk's own eight
    const int k0 = (x0 >> 5) - gw;               // plane word held by s_c[.][0]
    for (int i = threadIdx.x; i <= Rw; i += WIDE_CHUNK) s_h[i] = tab.h[i];
    if (threadIdx.x == 0) s_fl = 0;
    // all rows of the strip are staged at once (every thread has a dozen independent loads in flight); columns
    // outside the image hold no 0: "far" (255 > every h) / clear
    const int span = WIDE_CHUNK + 2 * Rw, cspan = WIDE_CHUNK / 32 + 2 * gw;
    for (int i = threadIdx.x; i < (y1 - y0) * span; i += WIDE_CHUNK) {
        const int r = i / span, c = i - r * span, xx = x0 - Rw + c;
        s_g[r][c] = (xx >= 0 && xx < W) ? g[v.pix_off + (size_t)(y0 + r) * W + xx] : 255;
    }
    for (int i = threadIdx.x; i < (y1 - y0) * cspan; i += WIDE_CHUNK) {
        const int r = i / cspan, c = i - r * cspan, kk = k0 + c;
        s_c[r][c] = (kk >= 0 && kk < wpr) ? clear[v.word_off + (size_t)(y0 + r) * wpr + kk] : FULL;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    const int k = (x0 >> 5) + (threadIdx.x >> 5);
    uint32_t fl = 0;
    for (int y = y0; y < y1; y++) {
        const int buf = y - y0;
        bool keep = false;
        if (x < W) {
            const uint32_t *cw = s_c[buf];
            const int xl = x - 32 * k0; // bit index of the pixel in the staged clear words
            if ((cw[xl >> 5] >> (xl & 31)) & 1u) {
                const int lo = xl - Rw, hi = xl + Rw;
                bool all = true;
                for (int w = lo >> 5; w <= (hi >> 5) && all; w++) {
                    uint32_t m = FULL;
                    if (w == (lo >> 5)) m &= FULL << (lo & 31);
                    if (w == (hi >> 5)) m &= FULL >> (31 - (hi & 31));
                    all = (~cw[w] & m) == 0u;
                }
                keep = all;
                if (!all) { // within reach of a 0: the test against the half heights, outward from the pixel
                    const uint8_t *c = s_g[buf] + Rw + threadIdx.x;
                    keep = true;
                    for (int dx = 1; keep && dx <= Rw; dx++) keep = (c[dx] > s_h[dx]) && (c[-dx] > s_h[dx]);
                }
            }
        }
        const uint32_t word = __ballot_sync(FULL, keep);
        if ((threadIdx.x & 31) == 0 && k < wpr) {
            const uint32_t vm = valid_mask(W, k);
            const uint32_t res = (invert ? ~word : word) & vm;
            out[v.word_off + (size_t)y * wpr + k] = res;
            fl |= (res ? 1u : 0u) | ((res ^ vm) ? 2u : 0u);
        }
    }
    if (fl) atomicOr(&s_fl, fl);
    __syncthreads();
    // one global atomic per CTA at most, none once the vignette's flags are complete
    if (threadIdx.x == 0 && s_fl && (*(volatile uint32_t *)(flags_out + img) & s_fl) != s_fl) atomicOr(flags_out + img, s_fl);
}

static int isqrt_i(int v)
{
    int r = 0;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}

extern "C" int maze_morph_pass_wide(const uint32_t *in, uint32_t *out, const maze_vignette_t *vig, int n_img,
                                    const int64_t *cta_off_a, long long n_cta_a, const int64_t *cta_off_b,
                                    long long n_cta_b, int t, int invert, const uint32_t *flags_in, uint32_t *flags_out,
                                    uint8_t *g_scratch, uint32_t *plane_scratch, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_cta_a <= 0 || n_cta_b <= 0) return MAZE_OK;
    if (in == out || !plane_scratch || plane_scratch == out || n_cta_a >= (1ll << 31) || n_cta_b >= (1ll << 31)) return MAZE_ERR_BADARG;
    WideTab tab;
    for (int i = 0; i <= WIDE_MAX_R; i++) tab.h[i] = 0;
    if (t >= 0) { // squared-distance threshold: closed disk d2 <= t
        if (t > WIDE_MAX_R * WIDE_MAX_R) return MAZE_ERR_BADARG;
        tab.R = tab.Rw = isqrt_i(t);
        tab.use_phantom = 1;
        for (int d = 0; d <= tab.Rw; d++) tab.h[d] = (uint8_t)isqrt_i(t - d * d);
    } else {      // registered footprint: half chords per |dy| -> half heights per |dx|
        int R, w[MAZE_MAX_DISK_RADIUS + 1], ph;
        if (t == -1 || !maze_pass_table(t, &R, w, &ph)) return MAZE_ERR_BADARG;
        tab.R = R;
        tab.Rw = w[0];
        tab.use_phantom = ph;
        for (int d = 0; d <= tab.Rw; d++) {
            int hh = 0;
            for (int dy = 0; dy <= R; dy++)
                if (w[dy] >= d) hh = dy;
            tab.h[d] = (uint8_t)hh;
        }
    }
    MAZE_CUDA(cudaMemsetAsync(flags_out, 0, sizeof(uint32_t) * (size_t)n_img, s), "wide flags");
    MAZE_KERNEL(KID_WIDE_VDIST, s,
                k_wide_vdist<<<(unsigned)n_cta_a, 256, 0, s>>>(in, vig, n_img, cta_off_a, tab.R, invert, tab.use_phantom,
                                                             flags_in, g_scratch, plane_scratch));
    MAZE_KERNEL(KID_WIDE_ROWS, s,
                k_wide_rows<<<(unsigned)n_cta_b, WIDE_CHUNK, 0, s>>>(g_scratch, plane_scratch, vig, n_img, cta_off_b, tab, invert, out,
                                                                   flags_out));
    return MAZE_OK;
}
