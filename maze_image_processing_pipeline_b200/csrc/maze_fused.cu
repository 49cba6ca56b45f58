// Vignette-resident fused stage: ONE CTA runs threshold -> up to four thresholded-EDT passes ->
// 8-connected labelling for one whole vignette with its bit planes and the union-find resident in
// shared memory, then streams out the mask bytes, the int32 label image and the final bit plane.
// sm_100a; sized for up to 227 KB of shared memory per CTA.
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:649 / :405      threshold / bool cast
//   maze_ipp/isotropic.py:35-36, 66-67        erosion / dilation compares on the EDT (incl. scipy's
//                                             phantom background pixel for uniform planes -- the CTA
//                                             sees the whole vignette, so the flags are exact)
//   maze_ipp/loki/pipeline.py:430-433         label(): raster-order labels, 8-connectivity
//
// HBM traffic per pixel: 1 B image read + 1 B mask + 4 B labels + 1/8 B bit plane written.
#include "maze_common.cuh"

struct FusedPass {
    int R;
    int invert;
    int w[MAZE_MAX_DISK_RADIUS + 1];
};
struct FusedParams {
    int t_int;
    int n_pass;
    FusedPass pass[4];
};

__device__ __forceinline__ uint32_t smem_plane_load(const uint32_t *plane, int H, int W, int wpr, int yy, int kk,
                                                    uint32_t inv, bool phantom)
{
    if (kk < 0 || kk >= wpr) return FULL;
    if (yy < 0 || yy >= H) return (phantom && yy == -1 && kk == 0) ? 0xfffffffeu : FULL;
    uint32_t v = plane[yy * wpr + kk] ^ inv;
    return v | ~valid_mask(W, kk);
}

// block-wide exclusive scan for any power-of-two block size <= 1024; s_warp: >= 33 ints
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

__device__ __forceinline__ int smem_find(const int *P, int n)
{
    int p = ld_volatile(P + n);
    while (p != n) {
        n = p;
        p = ld_volatile(P + n);
    }
    return n;
}

__device__ __forceinline__ void smem_union(int *P, int a, int b)
{
    while (true) {
        a = smem_find(P, a);
        b = smem_find(P, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(P + a, b);
        if (old == a) return;
        a = old;
    }
}

// id of the word run that contains bit b of word m (runs are numbered in raster order)
__device__ __forceinline__ int run_id(const int *RB, int w, uint32_t m, int b)
{
    uint32_t starts = m & ~(m << 1);
    uint32_t low = b == 31 ? FULL : ((2u << b) - 1u);
    return RB[w] + __popc(starts & low) - 1;
}

template <int T>
__global__ void __launch_bounds__(T) k_vignette_fused(const uint8_t *__restrict__ image,
                                                      const maze_vignette_t *__restrict__ vig,
                                                      const int32_t *__restrict__ img_list, FusedParams prm, int wcap,
                                                      uint32_t *__restrict__ bits_out, uint8_t *__restrict__ mask,
                                                      int32_t *__restrict__ labels, int32_t *__restrict__ n_labels,
                                                      int32_t *__restrict__ fallback)
{
    extern __shared__ uint32_t s_mem[];
    __shared__ int s_warp[34];
    const int img = img_list[blockIdx.x];
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, wpr = v.wpr, words = H * wpr;
    uint32_t *A = s_mem, *B = s_mem + wcap;
    int *P = (int *)(s_mem + 2 * wcap);
    const int tid = threadIdx.x;

    // ---- 1. threshold + pack (loki/pipeline.py:649) -------------------------------------------------
    {
        const uint8_t *base = image + v.pix_off;
        const int t = prm.t_int;
        const uint32_t t4 = (uint32_t)(t & 0xff) * 0x01010101u;
        for (int w = tid; w < words; w += T) {
            int y = w / wpr, k = w - y * wpr;
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
        for (int w = tid; w < words; w += T) {
            int k = w % wpr;
            has_zero |= ((src[w] ^ inv) | ~valid_mask(W, k)) != FULL;
        }
        const bool phantom = !__syncthreads_or(has_zero);
        for (int w = tid; w < words; w += T) {
            int y = w / wpr, k = w - y * wpr;
            uint32_t acc = FULL;
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
            dst[w] = (acc ^ inv) & valid_mask(W, k);
        }
        __syncthreads();
        uint32_t *tmp = src; src = dst; dst = tmp;
    }
    const uint32_t *M = src; // final plane
    int *RB = (int *)dst;    // free plane: first run id of every word

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
    if (n_runs > wcap) { // more runs than union-find slots: leave this vignette to the generic kernels
        if (tid == 0) { fallback[img] = 1; n_labels[img] = 0; }
        return;
    }
    for (int w = lo; w < hi; w++) {
        uint32_t m = M[w];
        RB[w] = run;
        run += __popc(m & ~(m << 1));
    }
    for (int r = tid; r < n_runs; r += T) P[r] = r;
    __syncthreads();
    for (int w = tid; w < words; w += T) {
        uint32_t m = M[w];
        if (!m) continue;
        int y = w / wpr, k = w - y * wpr;
        uint32_t prev = k > 0 ? M[w - 1] : 0u;
        uint32_t next = k + 1 < wpr ? M[w + 1] : 0u;
        if ((m & 1u) && (prev >> 31)) smem_union(P, RB[w], run_id(RB, w - 1, prev, 31));
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
            smem_union(P, run_id(RB, w, m, b), run_id(RB, w - wpr, uc, b));
        }
        while (need_ul) {
            int b = __ffs(need_ul) - 1;
            need_ul &= need_ul - 1;
            int tgt = b > 0 ? run_id(RB, w - wpr, uc, b - 1) : run_id(RB, w - wpr - 1, ulw, 31);
            smem_union(P, run_id(RB, w, m, b), tgt);
        }
        while (need_ur) {
            int b = __ffs(need_ur) - 1;
            need_ur &= need_ur - 1;
            int tgt = b < 31 ? run_id(RB, w - wpr, uc, b + 1) : RB[w - wpr + 1];
            smem_union(P, run_id(RB, w, m, b), tgt);
        }
    }
    __syncthreads();
    for (int r = tid; r < n_runs; r += T) {
        int root = smem_find(P, r);
        if (root != r) P[r] = root;
    }
    __syncthreads();
    // roots in raster order get labels 1..N, stored negated in their own slot
    const int rchunk = (n_runs + T - 1) / T;
    const int rlo = min(tid * rchunk, n_runs), rhi = min(rlo + rchunk, n_runs);
    int nroot = 0;
    for (int r = rlo; r < rhi; r++) nroot += (P[r] == r);
    int n_lab;
    int rank = block_exclusive_scan<T>(nroot, s_warp, &n_lab);
    __syncthreads();
    for (int r = rlo; r < rhi; r++)
        if (P[r] == r) P[r] = -(++rank);
    if (tid == 0) { n_labels[img] = n_lab; fallback[img] = 0; }

    // ---- 4. outputs: final bit plane, zero-filled mask / labels, then the foreground runs -------------
    {
        uint32_t *gb = bits_out + v.word_off;
        for (int w = tid; w < words; w += T) gb[w] = M[w];
        const int npx = H * W;
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *l4 = (uint4 *)(labels + v.pix_off);
        for (int i = tid; i < (npx + 3) / 4; i += T) l4[i] = z;
        uint4 *m4 = (uint4 *)(mask + v.pix_off);
        for (int i = tid; i < (npx + 15) / 16; i += T) m4[i] = z;
    }
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5;
        int32_t *gl = labels + v.pix_off;
        uint8_t *gm = mask + v.pix_off;
        for (int wb = warp * 32; wb < words; wb += T) {
            int w = wb + lane;
            uint32_t m = w < words ? M[w] : 0u;
            uint32_t nz = __ballot_sync(FULL, m != 0u);
            while (nz) {
                int j = __ffs(nz) - 1;
                nz &= nz - 1;
                uint32_t mj = __shfl_sync(FULL, m, j);
                if ((mj >> lane) & 1u) {
                    int wj = wb + j;
                    int y = wj / wpr, k = wj - y * wpr;
                    int p = P[run_id(RB, wj, mj, lane)];
                    int lab = p < 0 ? -p : -P[p];
                    int o = y * W + 32 * k + lane;
                    gl[o] = lab;
                    gm[o] = 1;
                }
            }
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

template <int T>
static int launch_class(int n, size_t smem, cudaStream_t s, const uint8_t *image, const maze_vignette_t *vig,
                        const int32_t *list, const FusedParams &prm, int wcap, uint32_t *bits, uint8_t *mask,
                        int32_t *labels, int32_t *n_labels, int32_t *fallback)
{
    if (n <= 0) return MAZE_OK;
    MAZE_CUDA(cudaFuncSetAttribute(k_vignette_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
              "fused smem attribute");
    MAZE_KERNEL(KID_VIGNETTE_FUSED, s, k_vignette_fused<T><<<n, T, smem, s>>>(image, vig, list, prm, wcap, bits, mask,
                                                                             labels, n_labels, fallback));
    return MAZE_OK;
}

extern "C" int maze_vignette_stage(const uint8_t *image, const maze_vignette_t *vig, const int32_t *img_list,
                                   const int32_t *class_off_host, int t_int, int n_pass, const int32_t *pass_t_host,
                                   const int32_t *pass_invert_host, uint32_t *bits, uint8_t *mask, int32_t *labels,
                                   int32_t *n_labels, int32_t *fallback, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pass < 0 || n_pass > 4) return MAZE_ERR_BADARG;
    FusedParams prm;
    prm.t_int = t_int;
    prm.n_pass = n_pass;
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
    const int caps[3] = {MAZE_FUSED_CAP0, MAZE_FUSED_CAP1, MAZE_FUSED_CAP2};
    int rc;
    for (int c = 0; c < 3; c++) {
        int n = class_off_host[c + 1] - class_off_host[c];
        const int32_t *list = img_list + class_off_host[c];
        size_t smem = (size_t)caps[c] * 12;
        if (c == 0)
            rc = launch_class<128>(n, smem, s, image, vig, list, prm, caps[c], bits, mask, labels, n_labels, fallback);
        else if (c == 1)
            rc = launch_class<256>(n, smem, s, image, vig, list, prm, caps[c], bits, mask, labels, n_labels, fallback);
        else
            rc = launch_class<1024>(n, smem, s, image, vig, list, prm, caps[c], bits, mask, labels, n_labels, fallback);
        if (rc != MAZE_OK) return rc;
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
