// Per-label regionprops reduction.  sm_100a.
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:589-594   FindRegions(labels, image): one RegionProperties per label
//   maze_ipp/loki/pipeline.py:653       ImageProperties(mask, image): the whole mask is one region
//   maze_ipp/loki/pipeline.py:604-625   the properties that are read (bbox, label, intensity image,
//                                       and the skimage RegionProperties subset of SURVEY.md 8 a9)
//
// Pass 1 accumulates EXACT integers per object: raw moments up to order 3 of (row, col), intensity
// sum / zero count / extrema and the bounding box.  Inside a 32-pixel word the lanes are cut into
// label segments; the geometric sums of a segment are closed forms of its end points and the
// intensity sums come from warp prefix sums, so each segment costs one set of atomics.
// The finishing kernel turns the raw moments into central moments with exact 128-bit integer
// arithmetic (one rounding when converting to float64), then normalised / Hu moments, inertia
// tensor, axes, eccentricity and orientation in float64.
#include <math.h>

#include "maze_common.cuh"

enum { A_N = 0, A_R, A_C, A_RR, A_RC, A_CC, A_RRR, A_RRC, A_RCC, A_CCC, A_V, A_Z };
enum { E_RMIN = 0, E_RMAX, E_CMIN, E_CMAX, E_VMIN, E_VMAX };

__global__ void k_props_init(u64 *acc, int32_t *ext, int n_obj)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_obj) return;
#pragma unroll
    for (int j = 0; j < MAZE_NACC; j++) acc[(i64)i * MAZE_NACC + j] = 0;
    int32_t *e = ext + (i64)i * MAZE_NEXT;
    e[E_RMIN] = 0x7fffffff; e[E_RMAX] = -1; e[E_CMIN] = 0x7fffffff; e[E_CMAX] = -1;
    e[E_VMIN] = 0x7fffffff; e[E_VMAX] = -1; e[6] = 0; e[7] = 0;
}

__device__ __forceinline__ u64 pow2sum(u64 m) { return m * (m + 1) * (2 * m + 1) / 6; }          // sum_{i<=m} i^2
__device__ __forceinline__ u64 pow3sum(u64 m) { u64 t = m * (m + 1) / 2; return t * t; }         // sum_{i<=m} i^3

__global__ void __launch_bounds__(MAZE_CTA) k_props_accumulate(const int32_t *__restrict__ labels,
                                                               const uint32_t *__restrict__ bits,
                                                               const uint8_t *__restrict__ image,
                                                               const maze_vignette_t *__restrict__ vig,
                                                               const maze_tile_t *__restrict__ tiles,
                                                               const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                               u64 *acc, int32_t *ext, const int32_t *__restrict__ skip_base)
{
    TileCtx c = load_tile(vig, tiles);
    if (skip_base && skip_base[c.img] >= 0) return; // rows of this vignette come from the staging arrays
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = c.v.w;
    int obj0 = lab_off[c.img];
    int nlab = lab_off[c.img + 1] - obj0;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    int lq[4] = {0, 0, 0, 0}; // labels of four words at a time: their loads are in flight together
    for (int i = 0; i < 32; i++, k++) {
        if (k == c.v.wpr) { k = 0; y++; }
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        if ((i & 3) == 0) {
            int yy = y, kk = k;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                lq[u] = 0;
                if (widx + u < c.nwords) {
                    const int xx = 32 * kk + lane;
                    if (labels) {
                        if (xx < W) lq[u] = labels[c.v.pix_off + (i64)yy * W + xx];
                    } else {
                        lq[u] = (__ldg(bits + c.v.word_off + widx + u) >> lane) & 1u;
                    }
                }
                if (++kk == c.v.wpr) { kk = 0; yy++; }
            }
        }
        int l = (i & 3) == 0 ? lq[0] : (i & 3) == 1 ? lq[1] : (i & 3) == 2 ? lq[2] : lq[3];
        if (l < 0 || l > nlab) l = 0;
        if (!__ballot_sync(FULL, l > 0)) continue;
        int v = (image && l > 0) ? (int)__ldg(image + c.v.pix_off + (i64)y * W + x) : 0;
        int lp = __shfl_up_sync(FULL, l, 1);
        bool start = lane == 0 || lp != l;
        uint32_t starts = __ballot_sync(FULL, start);
        uint32_t zeros = __ballot_sync(FULL, l > 0 && v == 0);
        uint32_t higher = lane == 31 ? 0u : (starts >> (lane + 1));
        int e = higher ? lane + __ffs(higher) - 1 : 31; // last lane of my segment
        int pv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULL, pv, d);
            if (lane >= d) pv += t;
        }
        int mn = v, mx = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o1 = __shfl_down_sync(FULL, mn, d);
            int o2 = __shfl_down_sync(FULL, mx, d);
            if (lane + d <= e) { mn = min(mn, o1); mx = max(mx, o2); }
        }
        int pe = __shfl_sync(FULL, pv, e);
        if (start && l > 0) {
            int row = obj0 + l - 1;
            if (row < n_obj_cap) {
                u64 a = (u64)(32 * k + lane), b = (u64)(32 * k + e), n = b - a + 1, r = (u64)y;
                u64 s1 = n * (a + b) / 2;
                u64 s2 = pow2sum(b) - (a ? pow2sum(a - 1) : 0);
                u64 s3 = pow3sum(b) - (a ? pow3sum(a - 1) : 0);
                u64 *A = acc + (i64)row * MAZE_NACC;
                atomicAdd(A + A_N, n);
                atomicAdd(A + A_R, n * r);
                atomicAdd(A + A_C, s1);
                atomicAdd(A + A_RR, n * r * r);
                atomicAdd(A + A_RC, r * s1);
                atomicAdd(A + A_CC, s2);
                atomicAdd(A + A_RRR, n * r * r * r);
                atomicAdd(A + A_RRC, r * r * s1);
                atomicAdd(A + A_RCC, r * s2);
                atomicAdd(A + A_CCC, s3);
                int32_t *E = ext + (i64)row * MAZE_NEXT;
                atomicMin(E + E_RMIN, y);
                atomicMax(E + E_RMAX, y);
                atomicMin(E + E_CMIN, (int)a);
                atomicMax(E + E_CMAX, (int)b);
                if (image) {
                    uint32_t seg = (e == 31 ? FULL : ((2u << e) - 1u)) & ~((1u << lane) - 1u);
                    atomicAdd(A + A_V, (u64)(pe - pv + v));
                    atomicAdd(A + A_Z, (u64)__popc(zeros & seg));
                    atomicMin(E + E_VMIN, mn);
                    atomicMax(E + E_VMAX, mx);
                }
            }
        }
    }
}

// Run-based variant for label images whose labels are constant along the word runs of `bits`
// (the output of maze_label, also after clear_border / remove_small_objects): one thread per word,
// only words with foreground do any work, the label is read once per run.
__global__ void __launch_bounds__(MAZE_CTA) k_props_runs(const int32_t *__restrict__ labels,
                                                         const uint32_t *__restrict__ bits,
                                                         const uint8_t *__restrict__ image,
                                                         const maze_vignette_t *__restrict__ vig,
                                                         const maze_tile_t *__restrict__ tiles,
                                                         const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                         u64 *acc, int32_t *ext, const int32_t *__restrict__ skip_base)
{
    TileCtx c = load_tile(vig, tiles);
    if (skip_base && skip_base[c.img] >= 0) return; // rows of this vignette come from the staging arrays
    int widx = c.word0 + threadIdx.x;
    if (widx >= c.nwords) return;
    uint32_t m = __ldg(bits + c.v.word_off + widx);
    if (!m) return;
    const int W = c.v.w;
    int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
    int obj0 = lab_off[c.img];
    int nlab = lab_off[c.img + 1] - obj0;
    const int32_t *L = labels + c.v.pix_off + (i64)y * W + 32 * k;
    const uint8_t *I = image ? image + c.v.pix_off + (i64)y * W + 32 * k : nullptr;
    uint32_t starts = m & ~(m << 1);
    while (starts) {
        int b0 = __ffs(starts) - 1;
        starts &= starts - 1;
        uint32_t run = ~(m >> b0);
        int len = run ? __ffs(run) - 1 : 32 - b0;
        int l = L[b0];
        if (l <= 0 || l > nlab) continue;
        int row = obj0 + l - 1;
        if (row >= n_obj_cap) continue;
        u64 a = (u64)(32 * k + b0), b = a + len - 1, n = (u64)len, r = (u64)y;
        u64 s1 = n * (a + b) / 2;
        u64 s2 = pow2sum(b) - (a ? pow2sum(a - 1) : 0);
        u64 s3 = pow3sum(b) - (a ? pow3sum(a - 1) : 0);
        u64 *A = acc + (i64)row * MAZE_NACC;
        atomicAdd(A + A_N, n);
        atomicAdd(A + A_R, n * r);
        atomicAdd(A + A_C, s1);
        atomicAdd(A + A_RR, n * r * r);
        atomicAdd(A + A_RC, r * s1);
        atomicAdd(A + A_CC, s2);
        atomicAdd(A + A_RRR, n * r * r * r);
        atomicAdd(A + A_RRC, r * r * s1);
        atomicAdd(A + A_RCC, r * s2);
        atomicAdd(A + A_CCC, s3);
        int32_t *E = ext + (i64)row * MAZE_NEXT;
        atomicMin(E + E_RMIN, y);
        atomicMax(E + E_RMAX, y);
        atomicMin(E + E_CMIN, (int)a);
        atomicMax(E + E_CMAX, (int)b);
        if (I) {
            int sv = 0, sz = 0, mn = 255, mx = 0;
            for (int j = 0; j < len; j++) {
                int v = (int)__ldg(I + b0 + j);
                sv += v; sz += (v == 0); mn = min(mn, v); mx = max(mx, v);
            }
            atomicAdd(A + A_V, (u64)sv);
            atomicAdd(A + A_Z, (u64)sz);
            atomicMin(E + E_VMIN, mn);
            atomicMax(E + E_VMAX, mx);
        }
    }
}

__global__ void __launch_bounds__(MAZE_CTA) k_props_runs_high(const int32_t *__restrict__ labels,
                                                              const uint32_t *__restrict__ bits,
                                                              const maze_vignette_t *__restrict__ vig,
                                                              const maze_tile_t *__restrict__ tiles,
                                                              const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                              double *table, const int32_t *__restrict__ skip_base)
{
    TileCtx c = load_tile(vig, tiles);
    if (skip_base && skip_base[c.img] >= 0) return;
    int widx = c.word0 + threadIdx.x;
    if (widx >= c.nwords) return;
    uint32_t m = __ldg(bits + c.v.word_off + widx);
    if (!m) return;
    const int W = c.v.w;
    int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
    int obj0 = lab_off[c.img];
    int nlab = lab_off[c.img + 1] - obj0;
    const int32_t *L = labels + c.v.pix_off + (i64)y * W + 32 * k;
    uint32_t starts = m & ~(m << 1);
    while (starts) {
        int b0 = __ffs(starts) - 1;
        starts &= starts - 1;
        uint32_t run = ~(m >> b0);
        int len = run ? __ffs(run) - 1 : 32 - b0;
        int l = L[b0];
        if (l <= 0 || l > nlab) continue;
        int row = obj0 + l - 1;
        if (row >= n_obj_cap) continue;
        double *f = table + (i64)row * MAZE_NFEAT;
        double dr = (double)y - f[MAZE_F_CENTROID];
        double cc = f[MAZE_F_CENTROID + 1];
        double s1 = 0, s2 = 0, s3 = 0;
        for (int j = 0; j < len; j++) {
            double dc = (double)(32 * k + b0 + j) - cc;
            s1 += dc; s2 += dc * dc; s3 += dc * dc * dc;
        }
        double dr2 = dr * dr, dr3 = dr2 * dr;
        atomicAdd(f + MAZE_F_MU + 1 * 4 + 3, dr * s3);
        atomicAdd(f + MAZE_F_MU + 2 * 4 + 2, dr2 * s2);
        atomicAdd(f + MAZE_F_MU + 3 * 4 + 1, dr3 * s1);
        atomicAdd(f + MAZE_F_MU + 2 * 4 + 3, dr2 * s3);
        atomicAdd(f + MAZE_F_MU + 3 * 4 + 2, dr3 * s2);
        atomicAdd(f + MAZE_F_MU + 3 * 4 + 3, dr3 * s3);
    }
}

__device__ __forceinline__ double i128_to_double(__int128 v)
{
    bool neg = v < 0;
    unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    u64 hi = (u64)(u >> 64), lo = (u64)u;
    double d = (double)hi * 18446744073709551616.0 + (double)lo;
    return neg ? -d : d;
}

__device__ __forceinline__ int find_image(const int32_t *lab_off, int n_img, int row)
{
    // largest i with lab_off[i] <= row and lab_off[i+1] > row
    int lo = 0, hi = n_img;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (lab_off[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ void finish_shape(double *f, const double mu[4][4], bool high_order)
{
    const double PI = 3.14159265358979323846;
    double nu[4][4];
    for (int p = 0; p < 4; p++)
        for (int q = 0; q < 4; q++) {
            bool have = (p + q <= 3) || high_order;
            nu[p][q] = (p + q >= 2 && have) ? mu[p][q] / pow(mu[0][0], (p + q) / 2.0 + 1.0) : nan("");
            f[MAZE_F_MU + p * 4 + q] = have ? mu[p][q] : nan("");
            f[MAZE_F_NU + p * 4 + q] = nu[p][q];
        }
    {
        double t0 = nu[3][0] + nu[1][2], t1 = nu[2][1] + nu[0][3];
        double q0 = t0 * t0, q1 = t1 * t1;
        double n4 = 4 * nu[1][1], s = nu[2][0] + nu[0][2], d = nu[2][0] - nu[0][2];
        double *hu = f + MAZE_F_HU;
        hu[0] = s;
        hu[1] = d * d + n4 * nu[1][1];
        hu[3] = q0 + q1;
        hu[5] = d * (q0 - q1) + n4 * t0 * t1;
        t0 *= q0 - 3 * q1;
        t1 *= 3 * q0 - q1;
        q0 = nu[3][0] - 3 * nu[1][2];
        q1 = 3 * nu[2][1] - nu[0][3];
        hu[2] = q0 * q0 + q1 * q1;
        hu[4] = q0 * t0 + q1 * t1;
        hu[6] = q1 * t0 - q0 * t1;
    }
    double a = mu[0][2] / mu[0][0], b = -mu[1][1] / mu[0][0], c = mu[2][0] / mu[0][0];
    f[MAZE_F_T00] = a; f[MAZE_F_T01] = b; f[MAZE_F_T11] = c;
    double tr = 0.5 * (a + c), df = 0.5 * (a - c);
    double rad = sqrt(df * df + b * b);
    double l1 = tr + rad, l2 = tr - rad;
    if (l1 < 0) l1 = 0;
    if (l2 < 0) l2 = 0;
    f[MAZE_F_EIG] = l1; f[MAZE_F_EIG + 1] = l2;
    f[MAZE_F_AXIS_MAJOR] = 4 * sqrt(l1);
    f[MAZE_F_AXIS_MINOR] = 4 * sqrt(l2);
    f[MAZE_F_ECC] = (l1 == 0) ? 0.0 : sqrt(1 - l2 / l1);
    if (a - c == 0) f[MAZE_F_ORIENT] = (b < 0) ? PI / 4 : -PI / 4;
    else f[MAZE_F_ORIENT] = 0.5 * atan2(-2 * b, c - a);
}

__device__ void finish_row(double *f, const u64 *A, const int32_t *E, int img, int label, int has_image, int high_order);

// stage 0: everything from the integer accumulators (high-order mu zeroed for pass 2 if requested)
// stage 1: re-derive nu / Hu once the float64 high-order central moments have been accumulated
__global__ void k_props_finish(const u64 *__restrict__ acc, const int32_t *__restrict__ ext,
                               const int32_t *__restrict__ lab_off, int n_img, int n_obj_cap, int has_image,
                               int high_order, int stage, const int32_t *__restrict__ acc_base, double *table)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    int total = min(lab_off[n_img], n_obj_cap);
    if (row >= total) return;
    double *f = table + (i64)row * MAZE_NFEAT;
    if (acc_base && acc_base[find_image(lab_off, n_img, row)] >= 0) return; // staged by the fused kernel
    if (stage == 1) {
        if (!(f[MAZE_F_AREA] > 0)) return;
        double mu[4][4];
        for (int p = 0; p < 4; p++)
            for (int q = 0; q < 4; q++) mu[p][q] = f[MAZE_F_MU + p * 4 + q];
        finish_shape(f, mu, true);
        return;
    }
    int img = find_image(lab_off, n_img, row);
    int label = row - lab_off[img] + 1;
    finish_row(f, acc + (i64)row * MAZE_NACC, ext + (i64)row * MAZE_NEXT, img, label, has_image, high_order);
}

// everything that derives from the integer accumulators of one object
__device__ void finish_row(double *f, const u64 *A, const int32_t *E, int img, int label, int has_image, int high_order)
{
    for (int j = 0; j < MAZE_NFEAT; j++) f[j] = nan("");
    f[MAZE_F_LABEL] = (double)label;
    f[MAZE_F_IMAGE] = (double)img;
    u64 n = A[A_N];
    f[MAZE_F_AREA] = (double)n;
    if (!n) return;
    f[MAZE_F_BBOX + 0] = E[E_RMIN]; f[MAZE_F_BBOX + 1] = E[E_CMIN];
    f[MAZE_F_BBOX + 2] = E[E_RMAX] + 1; f[MAZE_F_BBOX + 3] = E[E_CMAX] + 1;
    double dn = (double)n;
    f[MAZE_F_CENTROID] = (double)A[A_R] / dn;
    f[MAZE_F_CENTROID + 1] = (double)A[A_C] / dn;
    // exact central moments: n^(p+q-1) * mu_pq as 128-bit integers
    __int128 N = (__int128)n, R = (__int128)A[A_R], C = (__int128)A[A_C];
    __int128 RR = (__int128)A[A_RR], RC = (__int128)A[A_RC], CC = (__int128)A[A_CC];
    __int128 RRR = (__int128)A[A_RRR], RRC = (__int128)A[A_RRC], RCC = (__int128)A[A_RCC], CCC = (__int128)A[A_CCC];
    double mu[4][4];
    for (int p = 0; p < 4; p++)
        for (int q = 0; q < 4; q++) mu[p][q] = 0.0;
    double dn2 = dn * dn;
    mu[0][0] = dn;
    mu[2][0] = i128_to_double(N * RR - R * R) / dn;
    mu[1][1] = i128_to_double(N * RC - R * C) / dn;
    mu[0][2] = i128_to_double(N * CC - C * C) / dn;
    mu[3][0] = i128_to_double(N * N * RRR - 3 * N * R * RR + 2 * R * R * R) / dn2;
    mu[2][1] = i128_to_double(N * N * RRC - N * C * RR - 2 * N * R * RC + 2 * R * R * C) / dn2;
    mu[1][2] = i128_to_double(N * N * RCC - N * R * CC - 2 * N * C * RC + 2 * C * C * R) / dn2;
    mu[0][3] = i128_to_double(N * N * CCC - 3 * N * C * CC + 2 * C * C * C) / dn2;
    finish_shape(f, mu, false);
    if (high_order) {
        // pass 2 accumulates into these slots
        f[MAZE_F_MU + 1 * 4 + 3] = 0; f[MAZE_F_MU + 2 * 4 + 2] = 0; f[MAZE_F_MU + 3 * 4 + 1] = 0;
        f[MAZE_F_MU + 2 * 4 + 3] = 0; f[MAZE_F_MU + 3 * 4 + 2] = 0; f[MAZE_F_MU + 3 * 4 + 3] = 0;
    }
    if (has_image) {
        f[MAZE_F_IMIN] = E[E_VMIN];
        f[MAZE_F_IMAX] = E[E_VMAX];
        f[MAZE_F_IMEAN] = (double)A[A_V] / dn;
        f[MAZE_F_FRAC_INVALID] = (double)A[A_Z] / dn;
    }
}

// pass 2 (optional): float64 central moments with p + q > 3 about the exact centroid
__global__ void __launch_bounds__(MAZE_CTA) k_props_high_order(const int32_t *__restrict__ labels,
                                                               const uint32_t *__restrict__ bits,
                                                               const maze_vignette_t *__restrict__ vig,
                                                               const maze_tile_t *__restrict__ tiles,
                                                               const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                               double *table, const int32_t *__restrict__ skip_base)
{
    TileCtx c = load_tile(vig, tiles);
    if (skip_base && skip_base[c.img] >= 0) return;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = c.v.w;
    int obj0 = lab_off[c.img];
    int nlab = lab_off[c.img + 1] - obj0;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    int lq[4] = {0, 0, 0, 0}; // labels of four words at a time: their loads are in flight together
    for (int i = 0; i < 32; i++, k++) {
        if (k == c.v.wpr) { k = 0; y++; }
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        if ((i & 3) == 0) {
            int yy = y, kk = k;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                lq[u] = 0;
                if (widx + u < c.nwords) {
                    const int xx = 32 * kk + lane;
                    if (labels) {
                        if (xx < W) lq[u] = labels[c.v.pix_off + (i64)yy * W + xx];
                    } else {
                        lq[u] = (__ldg(bits + c.v.word_off + widx + u) >> lane) & 1u;
                    }
                }
                if (++kk == c.v.wpr) { kk = 0; yy++; }
            }
        }
        int l = (i & 3) == 0 ? lq[0] : (i & 3) == 1 ? lq[1] : (i & 3) == 2 ? lq[2] : lq[3];
        if (l < 0 || l > nlab) l = 0;
        if (!__ballot_sync(FULL, l > 0)) continue;
        int lp = __shfl_up_sync(FULL, l, 1);
        bool start = lane == 0 || lp != l;
        uint32_t starts = __ballot_sync(FULL, start);
        if (start && l > 0) {
            uint32_t higher = lane == 31 ? 0u : (starts >> (lane + 1));
            int e = higher ? lane + __ffs(higher) - 1 : 31;
            int row = obj0 + l - 1;
            if (row < n_obj_cap) {
                double *f = table + (i64)row * MAZE_NFEAT;
                double dr = (double)y - f[MAZE_F_CENTROID];
                double cc = f[MAZE_F_CENTROID + 1];
                double s1 = 0, s2 = 0, s3 = 0;
                for (int j = lane; j <= e; j++) {
                    double dc = (double)(32 * k + j) - cc;
                    s1 += dc; s2 += dc * dc; s3 += dc * dc * dc;
                }
                double dr2 = dr * dr, dr3 = dr2 * dr;
                atomicAdd(f + MAZE_F_MU + 1 * 4 + 3, dr * s3);
                atomicAdd(f + MAZE_F_MU + 2 * 4 + 2, dr2 * s2);
                atomicAdd(f + MAZE_F_MU + 3 * 4 + 1, dr3 * s1);
                atomicAdd(f + MAZE_F_MU + 2 * 4 + 3, dr2 * s3);
                atomicAdd(f + MAZE_F_MU + 3 * 4 + 2, dr3 * s2);
                atomicAdd(f + MAZE_F_MU + 3 * 4 + 3, dr3 * s3);
            }
        }
    }
}

extern "C" int maze_regionprops(const int32_t *labels, const uint32_t *bits, const uint8_t *image,
                                const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles, int n_tiles,
                                const int32_t *lab_off, int n_obj_cap, unsigned long long *acc, int32_t *ext,
                                double *table, int flags, const int32_t *acc_base, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0 || n_obj_cap <= 0) return MAZE_OK;
    if (!labels && !bits) return MAZE_ERR_BADARG;
    int nb = (n_obj_cap + 255) / 256;
    int high = (flags & MAZE_RP_HIGH_ORDER) ? 1 : 0;
    MAZE_KERNEL(KID_PROPS_INIT, s, k_props_init<<<nb, 256, 0, s>>>(acc, ext, n_obj_cap));
    const bool runs = (flags & MAZE_RP_RUNS) && labels && bits;
    if (runs)
        MAZE_KERNEL(KID_PROPS_RUNS, s, k_props_runs<<<n_tiles, MAZE_CTA, 0, s>>>(labels, bits, image, vig, tiles, lab_off, n_obj_cap, acc, ext, acc_base));
    else
        MAZE_KERNEL(KID_PROPS_ACCUMULATE, s, k_props_accumulate<<<n_tiles, MAZE_CTA, 0, s>>>(labels, bits, image, vig, tiles, lab_off, n_obj_cap, acc, ext, acc_base));
    MAZE_KERNEL(KID_PROPS_FINISH, s, k_props_finish<<<nb, 256, 0, s>>>(acc, ext, lab_off, n_img, n_obj_cap, image ? 1 : 0, high, 0, acc_base, table));
    if (high) {
        if (runs)
            MAZE_KERNEL(KID_PROPS_RUNS_HIGH, s, k_props_runs_high<<<n_tiles, MAZE_CTA, 0, s>>>(labels, bits, vig, tiles, lab_off, n_obj_cap, table, acc_base));
        else
            MAZE_KERNEL(KID_PROPS_HIGH_ORDER, s, k_props_high_order<<<n_tiles, MAZE_CTA, 0, s>>>(labels, bits, vig, tiles, lab_off, n_obj_cap, table, acc_base));
        MAZE_KERNEL(KID_PROPS_FINISH, s, k_props_finish<<<nb, 256, 0, s>>>(acc, ext, lab_off, n_img, n_obj_cap, image ? 1 : 0, high, 1, acc_base, table));
    }
    return MAZE_OK;
}


// Feature rows from the accumulators staged by k_vignette_fused (maze_fused.cu).
__global__ void k_props_finish_staged(const u64 *__restrict__ acc_stage, const double *__restrict__ hi_stage,
                                      const int32_t *__restrict__ ext_stage, const int32_t *__restrict__ acc_base,
                                      const int32_t *__restrict__ lab_off, int n_img, int n_obj, int has_image,
                                      int high_order, double *table)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= min(n_obj, lab_off[n_img])) return; // n_obj may be a capacity: the true total is on the device
    int img = find_image(lab_off, n_img, row);
    int base = acc_base[img];
    if (base < 0) return;
    int l = row - lab_off[img];
    i64 src = (i64)base + l;
    double *f = table + (i64)row * MAZE_NFEAT;
    finish_row(f, acc_stage + src * MAZE_NACC, ext_stage + src * MAZE_NEXT, img, l + 1, has_image, 0);
    if (high_order && f[MAZE_F_AREA] > 0) {
        const double *h = hi_stage + src * 8;
        double mu[4][4];
        for (int p = 0; p < 4; p++)
            for (int q = 0; q < 4; q++) mu[p][q] = f[MAZE_F_MU + p * 4 + q];
        mu[1][3] = h[0]; mu[2][2] = h[1]; mu[3][1] = h[2]; mu[2][3] = h[3]; mu[3][2] = h[4]; mu[3][3] = h[5];
        finish_shape(f, mu, true);
    }
}

extern "C" int maze_props_finish_staged(const unsigned long long *acc_stage, const double *hi_stage,
                                        const int32_t *ext_stage, const int32_t *acc_base, const int32_t *lab_off,
                                        int n_img, int n_obj, int has_intensity, int flags, double *table,
                                        void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_obj <= 0) return MAZE_OK;
    static thread_local int attr_dev = -1;
    int dev = 0;
    MAZE_CUDA(cudaGetDevice(&dev), "get device");
    if (attr_dev != dev) { // same carve-out as the stage kernels, so that it can share an SM with them
        MAZE_CUDA(cudaFuncSetAttribute(k_props_finish_staged, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "finish carveout");
        attr_dev = dev;
    }
    MAZE_KERNEL(KID_PROPS_FINISH_STAGED, s,
                k_props_finish_staged<<<(n_obj + 127) / 128, 128, 0, s>>>((const u64 *)acc_stage, hi_stage, ext_stage,
                                                                          acc_base, lab_off, n_img, n_obj, has_intensity,
                                                                          (flags & MAZE_RP_HIGH_ORDER) ? 1 : 0, table));
    return MAZE_OK;
}
