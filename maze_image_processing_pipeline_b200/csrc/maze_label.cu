// 8-connected component labelling on bit planes, label filters.  sm_100a.
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:430-433   skimage.measure.label(bool) == ndi.label(full 3x3): labels
//                                       1..N in raster order of each component's first pixel, int32
//   maze_ipp/loki/pipeline.py:435-439   clear_border(labels, out=labels)
//   maze_ipp/loki/pipeline.py:442-448   remove_small_objects(labels, min_size, out=labels)
//
// Union-find nodes are WORD RUNS: maximal runs of 1 bits inside one 32-bit word, identified by the
// per-vignette linear index of their first pixel.  Linking is lock-free with the smaller index as
// parent, so the root of a component is its first pixel in raster order and the final label is
// 1 + (number of roots before it in raster order) -- a prefix sum over root flags.
// The per-pixel parent array is only touched at run starts, so its DRAM traffic is a few sectors
// per run, not 4 B per pixel.
#include "maze_common.cuh"

// ---- 1. init: parent[start] = start for every word run ----------------------------------------
__global__ void __launch_bounds__(MAZE_CTA) k_ccl_init(const uint32_t *__restrict__ bits,
                                                       const maze_vignette_t *__restrict__ vig,
                                                       const maze_tile_t *__restrict__ tiles, int32_t *__restrict__ parent)
{
    TileCtx c = load_tile(vig, tiles);
    int widx = c.word0 + threadIdx.x;
    if (widx >= c.nwords) return;
    uint32_t m = __ldg(bits + c.v.word_off + widx);
    if (!m) return;
    int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
    int base = y * c.v.w + 32 * k;
    int32_t *P = parent + c.v.pix_off;
    uint32_t starts = m & ~(m << 1);
    while (starts) {
        int b = __ffs(starts) - 1;
        starts &= starts - 1;
        P[base + b] = base + b;
    }
}

// ---- 2. union: horizontal links between words, links to the row above ---------------------------
__global__ void __launch_bounds__(MAZE_CTA) k_ccl_union(const uint32_t *__restrict__ bits,
                                                        const maze_vignette_t *__restrict__ vig,
                                                        const maze_tile_t *__restrict__ tiles, int32_t *parent)
{
    TileCtx c = load_tile(vig, tiles);
    int widx = c.word0 + threadIdx.x;
    if (widx >= c.nwords) return;
    const uint32_t *plane = bits + c.v.word_off;
    uint32_t m = __ldg(plane + widx);
    if (!m) return;
    const int W = c.v.w, wpr = c.v.wpr;
    int y = widx / wpr, k = widx - y * wpr;
    int32_t *P = parent + c.v.pix_off;
    uint32_t prev = k > 0 ? __ldg(plane + widx - 1) : 0u;
    uint32_t next = k + 1 < wpr ? __ldg(plane + widx + 1) : 0u;
    int base = y * W + 32 * k;
    if ((m & 1u) && (prev >> 31)) uf_union(P, base, base - 32 + run_start_in_word(prev, 31));
    if (y == 0) return;
    uint32_t uc = __ldg(plane + widx - wpr);
    uint32_t ulw = k > 0 ? __ldg(plane + widx - wpr - 1) : 0u;
    uint32_t urw = k + 1 < wpr ? __ldg(plane + widx - wpr + 1) : 0u;
    if (!(uc | (ulw >> 31) | (urw & 1u))) return;
    uint32_t UL = (uc << 1) | (ulw >> 31);   // bit x = pixel (y-1, x-1)
    uint32_t UR = (uc >> 1) | (urw << 31);   // bit x = pixel (y-1, x+1)
    uint32_t left = (m << 1) | (prev >> 31); // bit x = pixel (y, x-1)
    uint32_t right = (m >> 1) | (next << 31);
    // one link per (run, upper run) contact is enough, see DESIGN.md "CCL"
    uint32_t need_up = m & uc & ~(left & UL);
    uint32_t need_ul = m & ~uc & UL & ~left;
    uint32_t need_ur = m & ~uc & UR & ~right;
    int ubase = base - W;
    while (need_up) {
        int b = __ffs(need_up) - 1;
        need_up &= need_up - 1;
        uf_union(P, base + run_start_in_word(m, b), ubase + run_start_in_word(uc, b));
    }
    while (need_ul) {
        int b = __ffs(need_ul) - 1;
        need_ul &= need_ul - 1;
        int tgt = b > 0 ? ubase + run_start_in_word(uc, b - 1) : ubase - 32 + run_start_in_word(ulw, 31);
        uf_union(P, base + run_start_in_word(m, b), tgt);
    }
    while (need_ur) {
        int b = __ffs(need_ur) - 1;
        need_ur &= need_ur - 1;
        int tgt = b < 31 ? ubase + run_start_in_word(uc, b + 1) : ubase + 32;
        uf_union(P, base + run_start_in_word(m, b), tgt);
    }
}

// ---- 3. flatten + count roots per tile ----------------------------------------------------------
__global__ void __launch_bounds__(MAZE_CTA) k_ccl_flatten(const uint32_t *__restrict__ bits,
                                                          const maze_vignette_t *__restrict__ vig,
                                                          const maze_tile_t *__restrict__ tiles, int32_t *parent,
                                                          int32_t *__restrict__ tile_scan)
{
    __shared__ int s_cnt;
    TileCtx c = load_tile(vig, tiles);
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int widx = c.word0 + threadIdx.x;
    int roots = 0;
    if (widx < c.nwords) {
        uint32_t m = __ldg(bits + c.v.word_off + widx);
        if (m) {
            int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
            int base = y * c.v.w + 32 * k;
            int32_t *P = parent + c.v.pix_off;
            uint32_t starts = m & ~(m << 1);
            while (starts) {
                int b = __ffs(starts) - 1;
                starts &= starts - 1;
                int n = base + b;
                int r = uf_find(P, n);
                if (r == n) roots++;
                else P[n] = r;
            }
        }
    }
    roots = __reduce_add_sync(FULL, roots);
    if ((threadIdx.x & 31) == 0 && roots) atomicAdd(&s_cnt, roots);
    __syncthreads();
    if (threadIdx.x == 0) tile_scan[blockIdx.x] = s_cnt;
}

// ---- 4. exclusive scan of the per-tile root counts (single CTA) + per-vignette offsets ----------
__global__ void __launch_bounds__(1024) k_tile_scan(int32_t *tile_scan, int n_tiles,
                                                    const maze_vignette_t *__restrict__ vig, int n_img,
                                                    int32_t *__restrict__ lab_off)
{
    __shared__ int s_part[1024];
    int t = threadIdx.x;
    int chunk = (n_tiles + 1023) / 1024;
    int lo = t * chunk, hi = min(lo + chunk, n_tiles);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += tile_scan[i];
    s_part[t] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int d = 1; d < 1024; d <<= 1) {
        int v = t >= d ? s_part[t - d] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    int run = s_part[t] - sum;
    for (int i = lo; i < hi; i++) {
        int v = tile_scan[i];
        tile_scan[i] = run;
        run += v;
    }
    if (t == 1023) tile_scan[n_tiles] = s_part[1023];
    __syncthreads();
    for (int i = t; i <= n_img; i += 1024) lab_off[i] = i < n_img ? tile_scan[vig[i].tile0] : tile_scan[n_tiles];
}

// ---- 5. roots get their raster rank + 1, stored in the label image at the root pixel --------------
__global__ void __launch_bounds__(MAZE_CTA) k_ccl_assign(const uint32_t *__restrict__ bits,
                                                         const maze_vignette_t *__restrict__ vig,
                                                         const maze_tile_t *__restrict__ tiles,
                                                         const int32_t *__restrict__ parent,
                                                         const int32_t *__restrict__ tile_scan, int32_t *labels)
{
    __shared__ int s_scan[16];
    TileCtx c = load_tile(vig, tiles);
    int widx = c.word0 + threadIdx.x;
    uint32_t rootmask = 0;
    int base = 0;
    if (widx < c.nwords) {
        uint32_t m = __ldg(bits + c.v.word_off + widx);
        if (m) {
            int y = widx / c.v.wpr, k = widx - y * c.v.wpr;
            base = y * c.v.w + 32 * k;
            const int32_t *P = parent + c.v.pix_off;
            uint32_t starts = m & ~(m << 1);
            while (starts) {
                int b = __ffs(starts) - 1;
                starts &= starts - 1;
                if (P[base + b] == base + b) rootmask |= 1u << b;
            }
        }
    }
    int total;
    int excl = cta_exclusive_scan(__popc(rootmask), s_scan, &total);
    if (!rootmask) return;
    int rank = tile_scan[blockIdx.x] - tile_scan[c.v.tile0] + excl;
    int32_t *L = labels + c.v.pix_off;
    while (rootmask) {
        int b = __ffs(rootmask) - 1;
        rootmask &= rootmask - 1;
        L[base + b] = ++rank;
    }
}

// ---- 6. final write: every pixel gets the label stored at its component's root pixel ------------
// (the root pixel itself is rewritten with the value it already holds, so reading roots while
// other warps write is benign)
__global__ void __launch_bounds__(MAZE_CTA) k_ccl_write(const uint32_t *__restrict__ bits,
                                                        const maze_vignette_t *__restrict__ vig,
                                                        const maze_tile_t *__restrict__ tiles,
                                                        const int32_t *__restrict__ parent, int32_t *labels)
{
    TileCtx c = load_tile(vig, tiles);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *P = parent + c.v.pix_off;
    int32_t *L = labels + c.v.pix_off;
    int wbase = c.word0 + warp * 32;
    if ((c.v.w & 3) == 0) {
        // rows start 16-byte aligned in the int32 image: a warp writes FOUR words (128 pixels) per step, one
        // 16-byte store per lane; background quads need no lookups at all
        for (int i = 0; i < 32; i += 4) {
            int widx = wbase + i + (lane >> 3);
            if (wbase + i >= c.nwords) break;
            bool ok = widx < c.nwords;
            uint32_t m = ok ? __ldg(bits + c.v.word_off + widx) : 0u;
            int y = ok ? widx / c.v.wpr : 0, k = ok ? widx - y * c.v.wpr : 0;
            int x = 32 * k + 4 * (lane & 7);
            if (!ok || x >= c.v.w) continue;
            int base = y * c.v.w + 32 * k;
            uint32_t nib = (m >> (4 * (lane & 7))) & 0xfu;
            int4 v = make_int4(0, 0, 0, 0);
            if (nib) {
                int lab[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int b = 4 * (lane & 7) + j;
                    if ((m >> b) & 1u) lab[j] = ld_volatile(L + P[base + run_start_in_word(m, b)]);
                }
                v = make_int4(lab[0], lab[1], lab[2], lab[3]);
            }
            *(int4 *)(L + base + 4 * (lane & 7)) = v;
        }
        return;
    }
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        uint32_t m = __ldg(bits + c.v.word_off + widx);
        int x = 32 * k + lane;
        int base = y * c.v.w + 32 * k;
        int lab = 0;
        if ((m >> lane) & 1u) {
            int n = base + run_start_in_word(m, lane);
            int r = P[n];
            lab = ld_volatile(L + r);
        }
        if (x < c.v.w) L[base + lane] = lab;
        if (++k == c.v.wpr) { k = 0; y++; }
    }
}

extern "C" int maze_label(const uint32_t *bits, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                          int n_tiles, int32_t *parent, int32_t *labels, int32_t *tile_scan, int32_t *lab_off,
                          void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    MAZE_KERNEL(KID_CCL_INIT, s, k_ccl_init<<<n_tiles, MAZE_CTA, 0, s>>>(bits, vig, tiles, parent));
    MAZE_KERNEL(KID_CCL_UNION, s, k_ccl_union<<<n_tiles, MAZE_CTA, 0, s>>>(bits, vig, tiles, parent));
    MAZE_KERNEL(KID_CCL_FLATTEN, s, k_ccl_flatten<<<n_tiles, MAZE_CTA, 0, s>>>(bits, vig, tiles, parent, tile_scan));
    MAZE_KERNEL(KID_TILE_SCAN, s, k_tile_scan<<<1, 1024, 0, s>>>(tile_scan, n_tiles, vig, n_img, lab_off));
    MAZE_KERNEL(KID_CCL_ASSIGN, s, k_ccl_assign<<<n_tiles, MAZE_CTA, 0, s>>>(bits, vig, tiles, parent, tile_scan, labels));
    MAZE_KERNEL(KID_CCL_WRITE, s, k_ccl_write<<<n_tiles, MAZE_CTA, 0, s>>>(bits, vig, tiles, parent, labels));
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------
// label filters
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int obj_row(const int32_t *lab_off, int img, int l, int n_obj_cap)
{
    int lo = lab_off[img], hi = lab_off[img + 1];
    if (l <= 0 || l > hi - lo) return -1;
    int r = lo + l - 1;
    return r < n_obj_cap ? r : -1;
}

__global__ void k_border_mark(const int32_t *__restrict__ labels, const maze_vignette_t *__restrict__ vig,
                              const int32_t *__restrict__ lab_off, int32_t *kill, int n_obj_cap)
{
    maze_vignette_t v = vig[blockIdx.x];
    const int32_t *L = labels + v.pix_off;
    int per = 2 * v.w + 2 * v.h;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
        int y, x;
        if (i < v.w) { y = 0; x = i; }
        else if (i < 2 * v.w) { y = v.h - 1; x = i - v.w; }
        else if (i < 2 * v.w + v.h) { y = i - 2 * v.w; x = 0; }
        else { y = i - 2 * v.w - v.h; x = v.w - 1; }
        int r = obj_row(lab_off, blockIdx.x, L[(i64)y * v.w + x], n_obj_cap);
        if (r >= 0) kill[r] = 1;
    }
}

// zero every pixel whose object is flagged: flag[row] != 0 (mode 0) or count[row] < min_size (mode 1)
__global__ void __launch_bounds__(MAZE_CTA) k_label_zero(int32_t *labels, const maze_vignette_t *__restrict__ vig,
                                                         const maze_tile_t *__restrict__ tiles,
                                                         const int32_t *__restrict__ lab_off,
                                                         const int32_t *__restrict__ obj, int n_obj_cap, int mode,
                                                         i64 min_size)
{
    TileCtx c = load_tile(vig, tiles);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *L = labels + c.v.pix_off;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        if (x < c.v.w) {
            i64 p = (i64)y * c.v.w + x;
            int l = L[p];
            int r = obj_row(lab_off, c.img, l, n_obj_cap);
            if (r >= 0) {
                bool kill = mode == 0 ? (obj[r] != 0) : ((i64)obj[r] < min_size);
                if (kill) L[p] = 0;
            }
        }
        if (++k == c.v.wpr) { k = 0; y++; }
    }
}

// pixel count per object; one atomic per label segment inside a 32-pixel word
__global__ void __launch_bounds__(MAZE_CTA) k_label_count(const int32_t *__restrict__ labels,
                                                          const maze_vignette_t *__restrict__ vig,
                                                          const maze_tile_t *__restrict__ tiles,
                                                          const int32_t *__restrict__ lab_off, int32_t *cnt,
                                                          int n_obj_cap)
{
    TileCtx c = load_tile(vig, tiles);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *L = labels + c.v.pix_off;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        int l = x < c.v.w ? L[(i64)y * c.v.w + x] : 0;
        if (__ballot_sync(FULL, l > 0)) {
            int lp = __shfl_up_sync(FULL, l, 1);
            bool start = lane == 0 || lp != l;
            uint32_t starts = __ballot_sync(FULL, start);
            if (start && l > 0) {
                uint32_t higher = lane == 31 ? 0u : (starts >> (lane + 1));
                int len = higher ? __ffs(higher) : 32 - lane;
                int r = obj_row(lab_off, c.img, l, n_obj_cap);
                if (r >= 0) atomicAdd(cnt + r, len);
            }
        }
        if (++k == c.v.wpr) { k = 0; y++; }
    }
}

extern "C" int maze_clear_border(int32_t *labels, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                                 int n_tiles, const int32_t *lab_off, int32_t *obj_scratch, int n_obj_cap,
                                 void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0 || n_obj_cap <= 0) return MAZE_OK;
    MAZE_CUDA(cudaMemsetAsync(obj_scratch, 0, sizeof(int32_t) * (size_t)n_obj_cap, s), "clear_border scratch");
    MAZE_KERNEL(KID_BORDER_MARK, s, k_border_mark<<<n_img, 256, 0, s>>>(labels, vig, lab_off, obj_scratch, n_obj_cap));
    MAZE_KERNEL(KID_LABEL_ZERO, s, k_label_zero<<<n_tiles, MAZE_CTA, 0, s>>>(labels, vig, tiles, lab_off, obj_scratch, n_obj_cap, 0, 0));
    return MAZE_OK;
}

extern "C" int maze_remove_small_objects(int32_t *labels, const maze_vignette_t *vig, int n_img,
                                         const maze_tile_t *tiles, int n_tiles, const int32_t *lab_off,
                                         int32_t *obj_scratch, int n_obj_cap, int64_t min_size, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0 || n_tiles <= 0 || n_obj_cap <= 0) return MAZE_OK;
    MAZE_CUDA(cudaMemsetAsync(obj_scratch, 0, sizeof(int32_t) * (size_t)n_obj_cap, s), "remove_small scratch");
    MAZE_KERNEL(KID_LABEL_COUNT, s, k_label_count<<<n_tiles, MAZE_CTA, 0, s>>>(labels, vig, tiles, lab_off, obj_scratch, n_obj_cap));
    MAZE_KERNEL(KID_LABEL_ZERO, s, k_label_zero<<<n_tiles, MAZE_CTA, 0, s>>>(labels, vig, tiles, lab_off, obj_scratch, n_obj_cap, 1, (i64)min_size));
    return MAZE_OK;
}

__global__ void __launch_bounds__(MAZE_CTA) k_max_label(const int32_t *__restrict__ labels,
                                                        const maze_vignette_t *__restrict__ vig,
                                                        const maze_tile_t *__restrict__ tiles, int32_t *max_label)
{
    TileCtx c = load_tile(vig, tiles);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *L = labels + c.v.pix_off;
    int wbase = c.word0 + warp * 32;
    int y = wbase / c.v.wpr, k = wbase - y * c.v.wpr;
    int mx = 0;
    for (int i = 0; i < 32; i++) {
        int widx = wbase + i;
        if (widx >= c.nwords) break;
        int x = 32 * k + lane;
        if (x < c.v.w) mx = max(mx, L[(i64)y * c.v.w + x]);
        if (++k == c.v.wpr) { k = 0; y++; }
    }
    mx = __reduce_max_sync(FULL, mx);
    if (lane == 0 && mx > 0) atomicMax(max_label + c.img, mx);
}

extern "C" int maze_max_label(const int32_t *labels, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                              int n_tiles, int32_t *max_label, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_img <= 0) return MAZE_OK;
    MAZE_CUDA(cudaMemsetAsync(max_label, 0, sizeof(int32_t) * (size_t)n_img, s), "max_label");
    if (n_tiles <= 0) return MAZE_OK;
    MAZE_KERNEL(KID_MAX_LABEL, s, k_max_label<<<n_tiles, MAZE_CTA, 0, s>>>(labels, vig, tiles, max_label));
    return MAZE_OK;
}


// threshold -> thresholded-EDT passes -> label -> mask bytes in ONE call (the per-operator chain for the
// vignettes the fused kernel cannot hold; saves a dozen host round trips through the binding)
extern "C" int maze_front_chain(const uint8_t *image, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                                int n_tiles, int t_int, int n_pass, const int32_t *pass_t_host,
                                const int32_t *pass_invert_host, uint32_t *plane_a, uint32_t *plane_b,
                                uint32_t *flags_a, uint32_t *flags_b, int32_t *parent, int32_t *labels,
                                int32_t *tile_scan, int32_t *lab_off, uint8_t *mask, uint32_t **final_plane_host,
                                void *stream)
{
    if (n_img <= 0 || n_tiles <= 0) return MAZE_OK;
    int rc = maze_threshold_pack(image, vig, n_img, tiles, n_tiles, t_int, plane_a, flags_a, stream);
    if (rc != MAZE_OK) return rc;
    uint32_t *src = plane_a, *dst = plane_b, *fs = flags_a, *fd = flags_b;
    for (int p = 0; p < n_pass; p++) {
        rc = maze_morph_pass(src, dst, vig, n_img, tiles, n_tiles, pass_t_host[p], pass_invert_host[p], fs, fd, stream);
        if (rc != MAZE_OK) return rc;
        uint32_t *t = src; src = dst; dst = t;
        t = fs; fs = fd; fd = t;
    }
    rc = maze_label(src, vig, n_img, tiles, n_tiles, parent, labels, tile_scan, lab_off, stream);
    if (rc != MAZE_OK) return rc;
    rc = maze_unpack_mask(src, vig, n_img, tiles, n_tiles, mask, stream);
    if (final_plane_host) *final_plane_host = src;
    return rc;
}
