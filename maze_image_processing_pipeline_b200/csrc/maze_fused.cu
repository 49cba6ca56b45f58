// Vignette-resident fused stage: ONE CTA runs threshold -> up to four thresholded-EDT passes ->
// 8-connected labelling -> per-label regionprops accumulation for one whole vignette, with its bit
// planes, the union-find and the per-label accumulators resident in shared memory, and streams out
// the mask bytes, the int32 label image, the final bit plane and one accumulator row per label.
// sm_100a; sized for up to 227 KB of shared memory per CTA (8 bytes per 32-pixel word).
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:649 / :405      threshold / bool cast
//   maze_ipp/isotropic.py:35-36, 66-67        erosion / dilation compares on the EDT (incl. scipy's
//                                             phantom background pixel for uniform planes -- the CTA
//                                             sees the whole vignette, so the flags are exact)
//   maze_ipp/loki/pipeline.py:430-433         label(): raster-order labels, 8-connectivity
//   maze_ipp/loki/pipeline.py:589-625         per-label RegionProperties reads (accumulators only;
//                                             k_props_finish_staged derives the features)
//
// HBM traffic per pixel: 1 B image read + 1 B mask + 4 B labels + 1/8 B bit plane written; the
// intensity re-read of foreground pixels hits L2 (the CTA read the vignette a moment earlier).
#include "maze_common.cuh"

#define FUSED_LCAP 24 /* labels per vignette whose accumulators live in shared memory */

enum { A_N = 0, A_R, A_C, A_RR, A_RC, A_CC, A_RRR, A_RRC, A_RCC, A_CCC, A_V, A_Z };
enum { E_RMIN = 0, E_RMAX, E_CMIN, E_CMAX, E_VMIN, E_VMAX };
enum { H_13 = 0, H_22, H_31, H_23, H_32, H_33, H_CR, H_CC };

struct FusedPass {
    int R;
    int invert;
    int w[MAZE_MAX_DISK_RADIUS + 1];
};
struct FusedParams {
    int t_int;
    int n_pass;
    int high_order;
    int stage_cap; // rows available in the staging arrays
    int do_props;  // 0: labels only, nothing is staged
    FusedPass pass[4];
};

struct AccRow { // shared-memory accumulator of one label
    u64 a[MAZE_NACC];
    double h[8];
    int e[MAZE_NEXT];
};

__device__ __forceinline__ uint32_t smem_plane_load(const uint32_t *plane, int H, int W, int wpr, int yy, int kk,
                                                    uint32_t inv, bool phantom)
{
    if (kk < 0 || kk >= wpr) return FULL;
    if (yy < 0 || yy >= H) return (phantom && yy == -1 && kk == 0) ? 0xfffffffeu : FULL;
    uint32_t v = plane[yy * wpr + kk] ^ inv;
    return v | ~valid_mask(W, kk);
}

// block-wide exclusive scan for any block size that is a multiple of 32 and <= 1024; s_warp: >= 33 ints
template <int T>
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int *total)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = T / 32;
        int s = lane < NW ? s_warp[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULL, si, d);
            if (lane >= d) si += t;
        }
        if (lane < NW) s_warp[lane] = si - s;
        if (lane == 31) s_warp[32] = si;
    }
    __syncthreads();
    int res = s_warp[warp] + inc - v;
    *total = s_warp[32];
    __syncthreads();
    return res;
}

// ---- union-find on 16-bit run ids in shared memory; values >= 0x8000 later encode "root, label" ----
typedef unsigned short u16;

__device__ __forceinline__ int find16(const u16 *P, int n)
{
    int p = *(const volatile u16 *)(P + n);
    while (p != n) {
        n = p;
        p = *(const volatile u16 *)(P + n);
    }
    return n;
}

// find with path halving (used while linking): a non-root's pointer may be replaced by any ancestor
__device__ __forceinline__ int find16_halve(u16 *P, int n)
{
    int p = *(volatile u16 *)(P + n);
    while (p != n) {
        int g = *(volatile u16 *)(P + p);
        if (g != p) *(volatile u16 *)(P + n) = (u16)g;
        n = p;
        p = g;
    }
    return n;
}

__device__ __forceinline__ void union16(u16 *P, int a, int b)
{
    while (true) {
        a = find16_halve(P, a);
        b = find16_halve(P, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicCAS(P + a, (u16)a, (u16)b); // a stays a root only while P[a] == a
        if (old == a) return;
        a = old;
    }
}

// id of the word run that contains bit b of word m (runs are numbered in raster order)
__device__ __forceinline__ int run_id(const u16 *RB, int w, uint32_t m, int b)
{
    uint32_t starts = m & ~(m << 1);
    uint32_t low = b == 31 ? FULL : ((2u << b) - 1u);
    return (int)RB[w] + __popc(starts & low) - 1;
}

__device__ __forceinline__ int label_of(const u16 *P, int rid)
{
    int p = P[rid];
    if (p < 0x8000) p = P[p];
    return p & 0x7fff;
}

// power sums of 0..m for m < 1024 (columns relative to a 1024-pixel chunk): 32-bit arithmetic suffices
__device__ __forceinline__ uint32_t f_pow2sum(uint32_t m) { return m * (m + 1) / 2 * (2 * m + 1) / 3; }
__device__ __forceinline__ u64 f_pow3sum(uint32_t m) { uint32_t t = m * (m + 1) / 2; return (u64)t * t; }

// One thresholded-EDT pass as a COLUMN WALK: a thread owns word column k and a strip of rows and slides
// down, keeping for the last 2R+1 rows the chord tests at the R+1 chord widths in registers, so every
// word is loaded once and every chord test is computed once (instead of 2R+1 times).
template <int R, int T>
__device__ __forceinline__ void morph_columns(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int H, int W,
                                              int wpr, const int *__restrict__ wtab, uint32_t inv, bool phantom,
                                              int nstrip, int S)
{
    int wd[R + 1];
#pragma unroll
    for (int j = 0; j <= R; j++) wd[j] = wtab[j];
    const int nitems = wpr * nstrip;
    for (int q = threadIdx.x; q < nitems; q += T) {
        const int s = q / wpr, k = q - s * wpr;
        const int y0 = s * S, y1 = min(H, y0 + S);
        if (y0 >= y1) continue;
        const uint32_t padC = ~valid_mask(W, k);
        const uint32_t padR = (k + 1 < wpr) ? ~valid_mask(W, k + 1) : 0u;
        uint32_t win[2 * R + 1][R + 1]; // win[i][j]: row (y - R + i) of the current output row y, chord width wd[j]
#pragma unroll
        for (int i = 0; i < 2 * R + 1; i++)
#pragma unroll
            for (int j = 0; j <= R; j++) win[i][j] = FULL;
        for (int r = y0 - R; r < y1 + R; r++) {
            uint32_t L = FULL, C = FULL, Rw = FULL;
            if (r >= 0 && r < H) {
                const uint32_t *row = src + r * wpr + k;
                C = (row[0] ^ inv) | padC;
                if (k > 0) L = row[-1] ^ inv;
                if (k + 1 < wpr) Rw = (row[1] ^ inv) | padR;
            } else if (phantom && r == -1) {
                if (k == 0) C = 0xfffffffeu;
                if (k == 1) L = 0xfffffffeu;
            }
            // slide the window up by one row
#pragma unroll
            for (int i = 0; i < 2 * R; i++)
#pragma unroll
                for (int j = 0; j <= R; j++) win[i][j] = win[i + 1][j];
            // chord tests of the new row, narrow to wide (wd[j] is non-increasing in j)
            uint32_t cur = C;
            int d = 0;
#pragma unroll
            for (int j = R; j >= 0; j--) {
                while (d < wd[j]) {
                    d++;
                    cur &= __funnelshift_rc(C, Rw, d) & __funnelshift_lc(L, C, d);
                }
                win[2 * R][j] = cur;
            }
            const int y = r - R;
            if (y >= y0) {
                uint32_t acc = FULL;
#pragma unroll
                for (int i = 0; i < 2 * R + 1; i++) acc &= win[i][i < R ? R - i : i - R];
                dst[y * wpr + k] = (acc ^ inv) & ~padC;
            }
        }
    }
}

template <int T>
__global__ void __launch_bounds__(T, 1024 / T) k_vignette_fused(
    const uint8_t *__restrict__ image, const uint8_t *__restrict__ intensity, const maze_vignette_t *__restrict__ vig,
    const int32_t *__restrict__ img_list, FusedParams prm, int wcap, uint32_t *__restrict__ bits_out,
    uint8_t *__restrict__ mask, int32_t *__restrict__ labels, int32_t *__restrict__ n_labels,
    int32_t *__restrict__ fallback, int32_t *__restrict__ acc_base, int32_t *stage_counter, u64 *acc_stage,
    double *hi_stage, int32_t *ext_stage)
{
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ int s_warp[34];
    __shared__ int s_base;
    const int img = img_list[blockIdx.x];
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, wpr = v.wpr, words = H * wpr;
    uint32_t *A = s_mem, *B = s_mem + wcap;
    AccRow *ACC = (AccRow *)(s_mem + 2 * wcap);
    const int tid = threadIdx.x;

    // ---- 0. zero-fill of the mask and label image: even CTAs issue it first (the stores drain while the
    // CTA computes), odd CTAs after the labelling, so the CTAs of a wave do not hit HBM in one burst
    auto zero_fill = [&]() {
        const int npx = H * W;
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *l4 = (uint4 *)(labels + v.pix_off);
        for (int i = tid; i < (npx + 3) / 4; i += T) l4[i] = z;
        uint4 *m4 = (uint4 *)(mask + v.pix_off);
        for (int i = tid; i < (npx + 15) / 16; i += T) m4[i] = z;
    };
    const bool fill_first = (blockIdx.x & 1) == 0;
    if (fill_first) zero_fill();
    // word walk without divisions: thread t visits words t, t + T, ...; (y, k) advance by (T / wpr, T % wpr)
    const int step_y = T / wpr, step_k = T - step_y * wpr;
    const int y_first = tid / wpr, k_first = tid - y_first * wpr;

    // ---- 1. threshold + pack (loki/pipeline.py:649) -------------------------------------------------
    {
        const uint8_t *base = image + v.pix_off;
        const int t = prm.t_int;
        const uint32_t t4 = (uint32_t)(t & 0xff) * 0x01010101u;
        int y = y_first, k = k_first;
        for (int w = tid; w < words; w += T) {
            int nvalid = min(32, W - 32 * k);
            const uint8_t *p = base + (size_t)y * W + 32 * k;
            uint32_t a = (uint32_t)((uintptr_t)p & 3u);
            const uint32_t *q = (const uint32_t *)(p - a);
            uint32_t word = 0;
            if (t < 0) {
                word = FULL;
            } else if (t < 255) {
                uint32_t lo = __ldg(q);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (4 * i < nvalid) {
                        uint32_t hi = __ldg(q + i + 1); // at most 4 bytes past the row: inside the padded slot
                        uint32_t px = __funnelshift_r(lo, hi, 8 * a);
                        uint32_t cmp = __vcmpgtu4(px, t4) & 0x01010101u;
                        word |= ((cmp * 0x01020408u) >> 24 & 0xfu) << (4 * i);
                        lo = hi;
                    }
                }
            }
            A[w] = word & valid_mask(W, k);
            y += step_y; k += step_k;
            if (k >= wpr) { k -= wpr; y++; }
        }
    }
    __syncthreads();

    // ---- 2. thresholded-EDT passes in shared memory (isotropic.py:35-36, 66-67) ----------------------
    uint32_t *src = A, *dst = B;
    for (int ps = 0; ps < prm.n_pass; ps++) {
        const int R = prm.pass[ps].R;
        const uint32_t inv = prm.pass[ps].invert ? FULL : 0u;
        // scipy's phantom background pixel: the (inverted) plane has no 0 at all
        bool has_zero = false;
        {
            int k = k_first;
            for (int w = tid; w < words; w += T) {
                has_zero |= ((src[w] ^ inv) | ~valid_mask(W, k)) != FULL;
                k += step_k;
                if (k >= wpr) k -= wpr;
            }
        }
        const bool phantom = !__syncthreads_or(has_zero);
        if (R >= 0 && R <= 3) {
            // strips of S rows x word columns: about one work item per thread
            const int nstrip = max(1, min(H, (T + wpr - 1) / wpr));
            const int S = (H + nstrip - 1) / nstrip;
            const int *wt = prm.pass[ps].w;
            if (R == 0) morph_columns<0, T>(src, dst, H, W, wpr, wt, inv, phantom, nstrip, S);
            else if (R == 1) morph_columns<1, T>(src, dst, H, W, wpr, wt, inv, phantom, nstrip, S);
            else if (R == 2) morph_columns<2, T>(src, dst, H, W, wpr, wt, inv, phantom, nstrip, S);
            else morph_columns<3, T>(src, dst, H, W, wpr, wt, inv, phantom, nstrip, S);
        } else {
            int y = y_first, k = k_first;
            for (int w = tid; w < words; w += T) {
                uint32_t acc = FULL;
                const bool interior = y >= R && y + R < H && k > 0 && (k + 2 < wpr || ((W & 31) == 0 && k + 1 < wpr));
                if (interior) { // no border, no padding bits, no phantom row in reach
                    for (int dy = -R; dy <= R; dy++) {
                        const uint32_t *row = src + w + dy * wpr;
                        int hw = prm.pass[ps].w[dy < 0 ? -dy : dy];
                        uint32_t C = row[0] ^ inv;
                        uint32_t h = C;
                        if (hw > 0) {
                            uint32_t L = row[-1] ^ inv, Rw = row[1] ^ inv;
                            for (int d = 1; d <= hw; d++) {
                                h &= __funnelshift_rc(C, Rw, d);
                                h &= __funnelshift_lc(L, C, d);
                            }
                        }
                        acc &= h;
                    }
                } else {
                    for (int dy = -R; dy <= R; dy++) {
                        int yy = y + dy;
                        int hw = prm.pass[ps].w[dy < 0 ? -dy : dy];
                        uint32_t C = smem_plane_load(src, H, W, wpr, yy, k, inv, phantom);
                        uint32_t h = C;
                        if (hw > 0) {
                            uint32_t L = smem_plane_load(src, H, W, wpr, yy, k - 1, inv, phantom);
                            uint32_t Rw = smem_plane_load(src, H, W, wpr, yy, k + 1, inv, phantom);
                            for (int d = 1; d <= hw; d++) {
                                h &= __funnelshift_rc(C, Rw, d);
                                h &= __funnelshift_lc(L, C, d);
                            }
                        }
                        acc &= h;
                    }
                }
                dst[w] = (acc ^ inv) & valid_mask(W, k);
                y += step_y; k += step_k;
                if (k >= wpr) { k -= wpr; y++; }
            }
        }
        __syncthreads();
        uint32_t *tmp = src; src = dst; dst = tmp;
    }
    const uint32_t *M = src;   // final plane
    u16 *RB = (u16 *)dst;      // free plane: first run id of every word ...
    u16 *P = RB + wcap;        // ... and the union-find over run ids

    // ---- 3. labelling on word runs in shared memory (loki/pipeline.py:430-433) -----------------------
    const int chunk = (words + T - 1) / T;
    const int lo = min(tid * chunk, words), hi = min(lo + chunk, words);
    int cnt = 0;
    for (int w = lo; w < hi; w++) {
        uint32_t m = M[w];
        cnt += __popc(m & ~(m << 1));
    }
    int n_runs;
    int run = block_exclusive_scan<T>(cnt, s_warp, &n_runs);
    if (n_runs > wcap || n_runs >= 0x8000) { // more runs than union-find slots: per-operator kernels take over
        if (tid == 0) { fallback[img] = 1; n_labels[img] = 0; acc_base[img] = -1; }
        return;
    }
    for (int w = lo; w < hi; w++) {
        uint32_t m = M[w];
        RB[w] = (u16)run;
        run += __popc(m & ~(m << 1));
    }
    for (int r = tid; r < n_runs; r += T) P[r] = (u16)r;
    __syncthreads();
    for (int w = tid; w < words; w += T) {
        uint32_t m = M[w];
        if (!m) continue;
        int y = w / wpr, k = w - y * wpr;
        uint32_t prev = k > 0 ? M[w - 1] : 0u;
        uint32_t next = k + 1 < wpr ? M[w + 1] : 0u;
        if ((m & 1u) && (prev >> 31)) union16(P, RB[w], run_id(RB, w - 1, prev, 31));
        if (y == 0) continue;
        uint32_t uc = M[w - wpr];
        uint32_t ulw = k > 0 ? M[w - wpr - 1] : 0u;
        uint32_t urw = k + 1 < wpr ? M[w - wpr + 1] : 0u;
        if (!(uc | (ulw >> 31) | (urw & 1u))) continue;
        uint32_t UL = (uc << 1) | (ulw >> 31), UR = (uc >> 1) | (urw << 31);
        uint32_t left = (m << 1) | (prev >> 31), right = (m >> 1) | (next << 31);
        uint32_t need_up = m & uc & ~(left & UL);
        uint32_t need_ul = m & ~uc & UL & ~left;
        uint32_t need_ur = m & ~uc & UR & ~right;
        while (need_up) {
            int b = __ffs(need_up) - 1;
            need_up &= need_up - 1;
            union16(P, run_id(RB, w, m, b), run_id(RB, w - wpr, uc, b));
        }
        while (need_ul) {
            int b = __ffs(need_ul) - 1;
            need_ul &= need_ul - 1;
            int tgt = b > 0 ? run_id(RB, w - wpr, uc, b - 1) : run_id(RB, w - wpr - 1, ulw, 31);
            union16(P, run_id(RB, w, m, b), tgt);
        }
        while (need_ur) {
            int b = __ffs(need_ur) - 1;
            need_ur &= need_ur - 1;
            int tgt = b < 31 ? run_id(RB, w - wpr, uc, b + 1) : (int)RB[w - wpr + 1];
            union16(P, run_id(RB, w, m, b), tgt);
        }
    }
    __syncthreads();
    for (int r = tid; r < n_runs; r += T) {
        int root = find16(P, r);
        if (root != r) P[r] = (u16)root;
    }
    __syncthreads();
    // roots in raster order get labels 1..N, stored in their own slot with the top bit set
    const int rchunk = (n_runs + T - 1) / T;
    const int rlo = min(tid * rchunk, n_runs), rhi = min(rlo + rchunk, n_runs);
    int nroot = 0;
    for (int r = rlo; r < rhi; r++) nroot += (P[r] == r);
    int n_lab;
    int rank = block_exclusive_scan<T>(nroot, s_warp, &n_lab);
    for (int r = rlo; r < rhi; r++)
        if (P[r] == r) P[r] = (u16)(0x8000 | (++rank));
    if (tid == 0) {
        n_labels[img] = n_lab;
        fallback[img] = 0;
        int base = -1;
        if (prm.do_props) {
            base = n_lab ? atomicAdd(stage_counter, n_lab) : 0;
            if (base + n_lab > prm.stage_cap) base = -1; // staging full: maze_regionprops takes this vignette
        }
        s_base = base;
        acc_base[img] = base;
    }

    // ---- 4. outputs: final bit plane, then the foreground runs over the zero-filled mask / labels -----
    if (!fill_first) zero_fill();
    {
        uint32_t *gb = bits_out + v.word_off;
        for (int w = tid; w < words; w += T) gb[w] = M[w];
        const int nsm = min(n_lab, FUSED_LCAP);
        for (int i = tid; i < nsm * (int)(sizeof(AccRow) / 4); i += T) ((uint32_t *)ACC)[i] = 0;
    }
    __syncthreads(); // labels in P are final; the zero-fill of step 0 is ordered before the stores below
    const int base = s_base;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = T / 32;
    {
        const int nsm = min(n_lab, FUSED_LCAP);
        for (int l = tid; l < nsm; l += T) {
            ACC[l].e[E_RMIN] = 0x7fffffff; ACC[l].e[E_RMAX] = -1; ACC[l].e[E_CMIN] = 0x7fffffff; ACC[l].e[E_CMAX] = -1;
            ACC[l].e[E_VMIN] = 0x7fffffff; ACC[l].e[E_VMAX] = -1;
        }
        if (base >= 0)
            for (int l = FUSED_LCAP + tid; l < n_lab; l += T) { // labels beyond the shared table accumulate in HBM
                u64 *ga = acc_stage + (i64)(base + l) * MAZE_NACC;
                for (int j = 0; j < MAZE_NACC; j++) ga[j] = 0;
                double *gh = hi_stage + (i64)(base + l) * 8;
                for (int j = 0; j < 8; j++) gh[j] = 0.0;
                int32_t *ge = ext_stage + (i64)(base + l) * MAZE_NEXT;
                ge[E_RMIN] = 0x7fffffff; ge[E_RMAX] = -1; ge[E_CMIN] = 0x7fffffff; ge[E_CMAX] = -1;
                ge[E_VMIN] = 0x7fffffff; ge[E_VMAX] = -1; ge[6] = 0; ge[7] = 0;
            }
        int32_t *gl = labels + v.pix_off;
        uint8_t *gm = mask + v.pix_off;
        // one warp per row, one lane per word: coalesced 4-byte label / 1-byte mask stores of set pixels
        for (int y = warp; y < H; y += NWARP) {
            for (int kc = 0; kc < wpr; kc += 32) {
                int k = kc + lane;
                uint32_t m = k < wpr ? M[y * wpr + k] : 0u;
                uint32_t nz = __ballot_sync(FULL, m != 0u);
                while (nz) {
                    int j = __ffs(nz) - 1;
                    nz &= nz - 1;
                    uint32_t mj = __shfl_sync(FULL, m, j);
                    if ((mj >> lane) & 1u) {
                        int o = y * W + 32 * (kc + j) + lane;
                        gl[o] = label_of(P, run_id(RB, y * wpr + kc + j, mj, lane));
                        gm[o] = 1;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (base < 0) return;

    // ---- 5. per-label accumulators (loki/pipeline.py:589-625) -----------------------------------------
    // COLUMN WALK: a thread owns word column k and a strip of rows.  Objects are vertically coherent, so
    // consecutive runs in a column almost always carry the same label: the sums of the current label stay
    // in registers and are flushed (shared-memory atomics) only when the label changes.
    const uint8_t *gi = intensity ? intensity + v.pix_off : nullptr;
    // short strips (<= 16 rows) and several work items per thread even out the foreground between threads
    const int p_S = max(1, min(16, (words + 2 * T - 1) / (2 * T)));
    const int p_nstrip = (H + p_S - 1) / p_S;
    const int p_items = wpr * p_nstrip;
    for (int q = tid; q < p_items; q += T) {
        const int st = q / wpr, k = q - st * wpr;
        const int y0 = st * p_S, y1 = min(H, y0 + p_S);
        int cur = 0;
        // moments of the current label in STRIP-LOCAL coordinates (row - y0 < 16, column - 32k < 32): 32 bits
        // are enough; the shift to vignette coordinates happens once per flush
        uint32_t aN = 0, aR = 0, aC = 0, aRR = 0, aRC = 0, aCC = 0, aRRR = 0, aRRC = 0, aRCC = 0, aCCC = 0;
        uint32_t aV = 0, aZ = 0;
        int rmin = 0, rmax = 0, cmin = 0x7fffffff, cmax = -1;
        uint32_t vmn = FULL, vmx = 0u;
        for (int y = y0; y <= y1; y++) {
            const uint32_t m = y < y1 ? M[y * wpr + k] : 0u;
            uint32_t pend = m;
            int rid = m ? (int)RB[y * wpr + k] : 0;
            const bool last = (y == y1);
            while (pend || last) {
                int L = 0, b0 = 0, len = 0;
                if (!last) {
                    b0 = __ffs(pend) - 1;
                    uint32_t rest = ~(pend >> b0);
                    len = rest ? __ffs(rest) - 1 : 32 - b0;
                    pend &= ~((len == 32 ? FULL : ((1u << len) - 1u)) << b0);
                    L = label_of(P, rid++);
                }
                if (L != cur) {
                    if (cur) { // flush the finished label: local -> vignette coordinates (r = y0 + r', c = cb + c')
                        const u64 N = aN, oy = (u64)y0, ox = 32 * (u64)k;
                        const u64 lR = aR, lC = aC, lRR = aRR, lRC = aRC, lCC = aCC;
                        const u64 gRR = oy * oy * N + 2 * oy * lR + lRR;
                        const u64 gCC = ox * ox * N + 2 * ox * lC + lCC;
                        u64 *Aa;
                        int *Ee;
                        if (cur <= FUSED_LCAP) { Aa = ACC[cur - 1].a; Ee = ACC[cur - 1].e; }
                        else { Aa = acc_stage + (i64)(base + cur - 1) * MAZE_NACC; Ee = ext_stage + (i64)(base + cur - 1) * MAZE_NEXT; }
                        atomicAdd(Aa + A_N, N);
                        atomicAdd(Aa + A_R, oy * N + lR);
                        atomicAdd(Aa + A_C, ox * N + lC);
                        atomicAdd(Aa + A_RR, gRR);
                        atomicAdd(Aa + A_RC, oy * ox * N + oy * lC + ox * lR + lRC);
                        atomicAdd(Aa + A_CC, gCC);
                        atomicAdd(Aa + A_RRR, oy * oy * oy * N + 3 * oy * oy * lR + 3 * oy * lRR + (u64)aRRR);
                        atomicAdd(Aa + A_RRC, ox * gRR + oy * oy * lC + 2 * oy * lRC + (u64)aRRC);
                        atomicAdd(Aa + A_RCC, oy * gCC + ox * ox * lR + 2 * ox * lRC + (u64)aRCC);
                        atomicAdd(Aa + A_CCC, ox * ox * ox * N + 3 * ox * ox * lC + 3 * ox * lCC + (u64)aCCC);
                        atomicMin(Ee + E_RMIN, rmin); atomicMax(Ee + E_RMAX, rmax);
                        atomicMin(Ee + E_CMIN, 32 * k + cmin); atomicMax(Ee + E_CMAX, 32 * k + cmax);
                        if (gi) {
                            atomicAdd(Aa + A_V, (u64)aV); atomicAdd(Aa + A_Z, (u64)aZ);
                            vmn = __vminu4(vmn, vmn >> 16); vmn = __vminu4(vmn, vmn >> 8);
                            vmx = __vmaxu4(vmx, vmx >> 16); vmx = __vmaxu4(vmx, vmx >> 8);
                            atomicMin(Ee + E_VMIN, (int)(vmn & 0xffu)); atomicMax(Ee + E_VMAX, (int)(vmx & 0xffu));
                        }
                    }
                    aN = aR = aC = aRR = aRC = aCC = aRRR = aRRC = aRCC = aCCC = 0;
                    aV = aZ = 0; cmin = 0x7fffffff; cmax = -1; vmn = FULL; vmx = 0u;
                    rmin = y;
                    cur = L;
                }
                if (last) break;
                // run [b0, b0 + len) of this word
                const uint32_t ja = b0, jb = b0 + len - 1, n = len, yl = (uint32_t)(y - y0);
                const uint32_t s1 = n * (ja + jb) / 2;
                const uint32_t s2 = f_pow2sum(jb) - (ja ? f_pow2sum(ja - 1) : 0u);
                const uint32_t t3b = jb * (jb + 1) / 2, t3a = ja ? (ja - 1) * ja / 2 : 0u;
                const uint32_t s3 = t3b * t3b - t3a * t3a;
                const uint32_t yn = yl * n, ys1 = yl * s1;
                aN += n; aR += yn; aC += s1; aRR += yl * yn; aRC += ys1; aCC += s2;
                aRRR += yl * yl * yn; aRRC += yl * ys1; aRCC += yl * s2; aCCC += s3;
                rmax = y;
                cmin = min(cmin, (int)ja); cmax = max(cmax, (int)jb);
                if (gi) { // intensity of the run's pixels, four at a time (byte-SIMD on aligned words)
                    const uint32_t runmask = (len == 32 ? FULL : ((1u << len) - 1u)) << b0;
                    const uint8_t *pw = gi + (size_t)y * W + 32 * (size_t)k;
                    const uint32_t al = (uint32_t)((uintptr_t)pw & 3u);
                    const uint32_t *qq = (const uint32_t *)(pw - al);
                    const int g0 = b0 >> 2, g1 = (b0 + len - 1) >> 2;
                    uint32_t lo32 = __ldg(qq + g0);
                    for (int g = g0; g <= g1; g++) {
                        uint32_t hi32 = __ldg(qq + g + 1); // <= 4 bytes past the row: inside the padded slot
                        uint32_t px = __funnelshift_r(lo32, hi32, 8 * al);
                        lo32 = hi32;
                        uint32_t nib = (runmask >> (4 * g)) & 0xfu;
                        uint32_t bm = ((nib * 0x00204081u) & 0x01010101u) * 0xffu;
                        aV += __vsadu4(px & bm, 0u);
                        aZ += __popc(__vcmpeq4(px, 0u) & bm & 0x01010101u);
                        vmn = __vminu4(vmn, px | ~bm);
                        vmx = __vmaxu4(vmx, px & bm);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (prm.high_order) {
        // float64 central moments with p + q > 3 about the exact centroid: sum_c (c - cc)^q over a run
        // follows from its n, S1, S2, S3; same column walk, six doubles per current label in registers
        for (int l = tid; l < n_lab; l += T) {
            const u64 *Aa = l < FUSED_LCAP ? ACC[l].a : acc_stage + (i64)(base + l) * MAZE_NACC;
            double *Hh = l < FUSED_LCAP ? ACC[l].h : hi_stage + (i64)(base + l) * 8;
            double dn = (double)*(const volatile u64 *)(Aa + A_N);
            Hh[H_CR] = (double)*(const volatile u64 *)(Aa + A_R) / dn;
            Hh[H_CC] = (double)*(const volatile u64 *)(Aa + A_C) / dn;
        }
        __syncthreads();
        for (int q = tid; q < p_items; q += T) {
            const int st = q / wpr, k = q - st * wpr;
            const int y0 = st * p_S, y1 = min(H, y0 + p_S);
            int cur = 0;
            double h13 = 0, h22 = 0, h31 = 0, h23 = 0, h32 = 0, h33 = 0, cr = 0, o = 0;
            double *Hh = nullptr;
            for (int y = y0; y <= y1; y++) {
                const uint32_t m = y < y1 ? M[y * wpr + k] : 0u;
                uint32_t pend = m;
                int rid = m ? (int)RB[y * wpr + k] : 0;
                bool last = (y == y1);
                while (pend || last) {
                    int L = 0, b0 = 0, len = 0;
                    if (!last) {
                        b0 = __ffs(pend) - 1;
                        uint32_t rest = ~(pend >> b0);
                        len = rest ? __ffs(rest) - 1 : 32 - b0;
                        pend &= ~((len == 32 ? FULL : ((1u << len) - 1u)) << b0);
                        L = label_of(P, rid++);
                    }
                    if (L != cur) {
                        if (cur) {
                            atomicAdd(Hh + H_13, h13); atomicAdd(Hh + H_22, h22); atomicAdd(Hh + H_31, h31);
                            atomicAdd(Hh + H_23, h23); atomicAdd(Hh + H_32, h32); atomicAdd(Hh + H_33, h33);
                        }
                        h13 = h22 = h31 = h23 = h32 = h33 = 0;
                        cur = L;
                        if (L) {
                            Hh = L <= FUSED_LCAP ? ACC[L - 1].h : hi_stage + (i64)(base + L - 1) * 8;
                            cr = *(volatile double *)(Hh + H_CR);
                            o = *(volatile double *)(Hh + H_CC) - 32.0 * k; // centroid column relative to the word
                        }
                    }
                    if (last) break;
                    const uint32_t ja = b0, jb = b0 + len - 1, n = len;
                    const uint32_t s1 = n * (ja + jb) / 2;
                    const uint32_t s2 = f_pow2sum(jb) - (ja ? f_pow2sum(ja - 1) : 0u);
                    const uint32_t t3b = jb * (jb + 1) / 2, t3a = ja ? (ja - 1) * ja / 2 : 0u;
                    const uint32_t s3 = t3b * t3b - t3a * t3a;
                    const double dn = (double)n, S1 = (double)s1, S2 = (double)s2, S3 = (double)s3;
                    const double T1 = S1 - dn * o;
                    const double T2 = S2 - 2.0 * o * S1 + dn * o * o;
                    const double T3 = S3 - 3.0 * o * S2 + 3.0 * o * o * S1 - dn * o * o * o;
                    const double dr = (double)y - cr, dr2 = dr * dr, dr3 = dr2 * dr;
                    h13 += dr * T3; h22 += dr2 * T2; h31 += dr3 * T1;
                    h23 += dr2 * T3; h32 += dr3 * T2; h33 += dr3 * T3;
                }
            }
        }
        __syncthreads();
    }
    // shared rows -> staging
    {
        const int nsm = min(n_lab, FUSED_LCAP);
        for (int i = tid; i < nsm * MAZE_NACC; i += T) {
            int l = i / MAZE_NACC, j = i - l * MAZE_NACC;
            acc_stage[(i64)(base + l) * MAZE_NACC + j] = ACC[l].a[j];
        }
        for (int i = tid; i < nsm * 8; i += T) {
            int l = i >> 3, j = i & 7;
            hi_stage[(i64)(base + l) * 8 + j] = ACC[l].h[j];
            ext_stage[(i64)(base + l) * MAZE_NEXT + j] = ACC[l].e[j];
        }
    }
}

__global__ void __launch_bounds__(1024) k_count_scan(const int32_t *__restrict__ n_labels, int n_img,
                                                     int32_t *__restrict__ lab_off)
{
    __shared__ int s_warp[34];
    int t = threadIdx.x;
    int chunk = (n_img + 1023) / 1024;
    int lo = min(t * chunk, n_img), hi = min(lo + chunk, n_img);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += n_labels[i];
    int total;
    int run = block_exclusive_scan<1024>(sum, s_warp, &total);
    for (int i = lo; i < hi; i++) {
        lab_off[i] = run;
        run += n_labels[i];
    }
    if (t == 0) lab_off[n_img] = total;
}

static int isqrt_i(int v)
{
    int r = 0;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}

struct FusedArgs {
    const uint8_t *image, *intensity;
    const maze_vignette_t *vig;
    uint32_t *bits;
    uint8_t *mask;
    int32_t *labels, *n_labels, *fallback, *acc_base, *stage_counter;
    u64 *acc_stage;
    double *hi_stage;
    int32_t *ext_stage;
};

// Forked streams for the concurrent class launches (per host thread and device; created on first use).
struct ForkStreams {
    int device;
    cudaStream_t aux[MAZE_FUSED_CLASSES];
    cudaEvent_t fork, join[MAZE_FUSED_CLASSES];
};

static ForkStreams *fork_streams()
{
    static thread_local ForkStreams pool[16];
    static thread_local int n_pool = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < n_pool; i++)
        if (pool[i].device == dev) return &pool[i];
    if (n_pool >= 16) return nullptr;
    ForkStreams *f = &pool[n_pool];
    f->device = dev;
    if (cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int c = 0; c < MAZE_FUSED_CLASSES; c++) {
        if (cudaStreamCreateWithFlags(&f->aux[c], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f->join[c], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    n_pool++;
    return f;
}

template <int T>
static int launch_class(int n, int wcap, cudaStream_t s, const int32_t *list, const FusedParams &prm,
                        const FusedArgs &a)
{
    if (n <= 0) return MAZE_OK;
    size_t smem = (size_t)wcap * 8 + FUSED_LCAP * sizeof(AccRow);
    MAZE_CUDA(cudaFuncSetAttribute(k_vignette_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
              "fused smem attribute");
    MAZE_CUDA(cudaFuncSetAttribute(k_vignette_fused<T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared), "fused carveout");
    MAZE_KERNEL(KID_VIGNETTE_FUSED, s,
                k_vignette_fused<T><<<n, T, smem, s>>>(a.image, a.intensity, a.vig, list, prm, wcap, a.bits, a.mask,
                                                       a.labels, a.n_labels, a.fallback, a.acc_base, a.stage_counter,
                                                       a.acc_stage, a.hi_stage, a.ext_stage));
    return MAZE_OK;
}

extern "C" int maze_vignette_stage(const uint8_t *image, const uint8_t *intensity, const maze_vignette_t *vig,
                                   const int32_t *img_list, const int32_t *class_off_host, int t_int, int n_pass,
                                   const int32_t *pass_t_host, const int32_t *pass_invert_host, int flags,
                                   uint32_t *bits, uint8_t *mask, int32_t *labels, int32_t *n_labels,
                                   int32_t *fallback, int32_t *acc_base, int32_t *stage_counter, int stage_cap,
                                   unsigned long long *acc_stage, double *hi_stage, int32_t *ext_stage, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pass < 0 || n_pass > 4) return MAZE_ERR_BADARG;
    FusedParams prm;
    prm.t_int = t_int;
    prm.n_pass = n_pass;
    prm.high_order = (flags & MAZE_RP_HIGH_ORDER) ? 1 : 0;
    prm.stage_cap = stage_cap;
    prm.do_props = (flags & MAZE_FUSED_NO_PROPS) ? 0 : 1;
    for (int p = 0; p < 4; p++) {
        prm.pass[p].R = -1;
        prm.pass[p].invert = 0;
        for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) prm.pass[p].w[i] = 0;
    }
    for (int p = 0; p < n_pass; p++) {
        int t = pass_t_host[p];
        if (t >= (MAZE_MAX_DISK_RADIUS + 1) * (MAZE_MAX_DISK_RADIUS + 1)) return MAZE_ERR_BADARG;
        prm.pass[p].invert = pass_invert_host[p] ? 1 : 0;
        if (t >= 0) {
            prm.pass[p].R = isqrt_i(t);
            for (int dy = 0; dy <= prm.pass[p].R; dy++) prm.pass[p].w[dy] = isqrt_i(t - dy * dy);
        }
    }
    MAZE_CUDA(cudaMemsetAsync(stage_counter, 0, sizeof(int32_t), s), "stage counter");
    FusedArgs a = {image, intensity, vig, bits, mask, labels, n_labels, fallback, acc_base, stage_counter,
                   (u64 *)acc_stage, hi_stage, ext_stage};
    const int caps[MAZE_FUSED_CLASSES] = {MAZE_FUSED_CAP0, MAZE_FUSED_CAP1, MAZE_FUSED_CAP2, MAZE_FUSED_CAP3};
    // The size classes touch disjoint vignettes, so their kernels run CONCURRENTLY: the class of the
    // largest vignettes (one long CTA per SM) goes to the caller's stream, the others to forked streams,
    // and the small CTAs fill the SMs the big ones leave idle.
    ForkStreams *fk = fork_streams();
    if (!fk) return MAZE_ERR_CUDA;
    MAZE_CUDA(cudaEventRecord(fk->fork, s), "fork record");
    for (int c = MAZE_FUSED_CLASSES - 1; c >= 0; c--) {
        int n = class_off_host[c + 1] - class_off_host[c];
        if (n <= 0) continue;
        const int32_t *list = img_list + class_off_host[c];
        cudaStream_t sc = s;
        if (c != MAZE_FUSED_CLASSES - 1) {
            sc = fk->aux[c];
            MAZE_CUDA(cudaStreamWaitEvent(sc, fk->fork, 0), "fork wait");
        }
        int rc;
        if (c == 0) rc = launch_class<128>(n, caps[c], sc, list, prm, a);
        else if (c == 1) rc = launch_class<256>(n, caps[c], sc, list, prm, a);
        else if (c == 2) rc = launch_class<512>(n, caps[c], sc, list, prm, a);
        else rc = launch_class<1024>(n, caps[c], sc, list, prm, a);
        if (rc != MAZE_OK) return rc;
        if (sc != s) {
            MAZE_CUDA(cudaEventRecord(fk->join[c], sc), "join record");
            MAZE_CUDA(cudaStreamWaitEvent(s, fk->join[c], 0), "join wait");
        }
    }
    return MAZE_OK;
}

extern "C" int maze_count_scan(const int32_t *n_labels, int n_img, int32_t *lab_off, void *stream)
{
    if (n_img < 0) return MAZE_ERR_BADARG;
    MAZE_KERNEL(KID_COUNT_SCAN, (cudaStream_t)stream,
                k_count_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(n_labels, n_img, lab_off));
    return MAZE_OK;
}
