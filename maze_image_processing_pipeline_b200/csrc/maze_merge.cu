// merge_labels on the device: one CTA per vignette runs the reference's whole sequential loop.
//
// Reference behaviour restated: maze_ipp/merge_labels.py:29-113 (helpers :7-26), called aliased
// (labels_out is labels) from maze_ipp/loki/pipeline.py:451-457.
//
// Distances are kept as exact integer squared distances (int32); the reference's float64 values
// are sqrt(d2), which is correctly rounded on the device, so `distmap + cur_distmap`,
// `sum.min()`, `> max_distance` and `<= merge_dist + path_tolerance` are evaluated on the very same
// float64 numbers, and `cur_distmap < distmap` / the per-label minima are decided on the integers
// (sqrt is strictly monotone on them).
#include <math.h>

#include "maze_common.cuh"

#define MG_CTA 1024
#define MG_INF (1 << 24)

struct MgShared {
    int red_i[MG_CTA / 32][4];
    double red_d[MG_CTA / 32];
    u64 red_u[MG_CTA / 32];
    int bbox[4];
    int ival;
    u64 uval;
    double dval;
};

__device__ __forceinline__ void mg_bbox(MgShared &S, const int32_t *L, int H, int W, int l, int *out /*r0,r1,c0,c1 inclusive*/)
{
    int r0 = 0x7fffffff, r1 = -1, c0 = 0x7fffffff, c1 = -1;
    int n = H * W;
    {
        const int sy = MG_CTA / W, sx = MG_CTA - sy * W;
        int y = threadIdx.x / W, x = threadIdx.x - y * W;
        for (int p0 = threadIdx.x; p0 < n; p0 += 4 * MG_CTA) { // four independent loads in flight
            int v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (p0 + u * MG_CTA < n) ? L[p0 + u * MG_CTA] : l - 1;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (v[u] == l) { r0 = min(r0, y); r1 = max(r1, y); c0 = min(c0, x); c1 = max(c1, x); }
                y += sy; x += sx;
                if (x >= W) { x -= W; y++; }
            }
        }
    }
    r0 = __reduce_min_sync(FULL, r0); r1 = __reduce_max_sync(FULL, r1);
    c0 = __reduce_min_sync(FULL, c0); c1 = __reduce_max_sync(FULL, c1);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { S.red_i[warp][0] = r0; S.red_i[warp][1] = r1; S.red_i[warp][2] = c0; S.red_i[warp][3] = c1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) {
            r0 = min(r0, S.red_i[w][0]); r1 = max(r1, S.red_i[w][1]);
            c0 = min(c0, S.red_i[w][2]); c1 = max(c1, S.red_i[w][3]);
        }
        S.bbox[0] = r0; S.bbox[1] = r1; S.bbox[2] = c0; S.bbox[3] = c1;
    }
    __syncthreads();
    for (int j = 0; j < 4; j++) out[j] = S.bbox[j];
    __syncthreads();
}

__device__ __forceinline__ int mg_max_int(MgShared &S, int v)
{
    v = __reduce_max_sync(FULL, v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_i[warp][0] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = max(v, S.red_i[w][0]);
        S.ival = v;
    }
    __syncthreads();
    v = S.ival;
    __syncthreads();
    return v;
}

__device__ __forceinline__ u64 mg_min_u64(MgShared &S, u64 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_u[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = S.red_u[w] < v ? S.red_u[w] : v;
        S.uval = v;
    }
    __syncthreads();
    v = S.uval;
    __syncthreads();
    return v;
}

__device__ __forceinline__ double mg_min_double(MgShared &S, double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_d[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = S.red_d[w] < v ? S.red_d[w] : v;
        S.dval = v;
    }
    __syncthreads();
    v = S.dval;
    __syncthreads();
    return v;
}

// merge_labels.py:12-26.  Squared distances to the pixels of label l for the window
// (bbox enlarged by pad, clipped) into D; *fill receives the window maximum, win the window
// (half open).  Returns false where the reference raises TypeError (label absent, :19-20).
__device__ bool mg_windowed_d2(MgShared &S, const int32_t *L, int H, int W, int l, int have_max, int pad,
                               int32_t *G, int32_t *D, int *win, int *fill)
{
    int bb[4];
    mg_bbox(S, L, H, W, l, bb);
    bool empty = bb[1] < 0;
    if (empty && have_max) return false;
    int r0 = 0, r1 = H, c0 = 0, c1 = W;
    if (have_max) {
        r0 = max(0, bb[0] - pad); r1 = (int)min((i64)H, (i64)bb[1] + 1 + pad);
        c0 = max(0, bb[2] - pad); c1 = (int)min((i64)W, (i64)bb[3] + 1 + pad);
    }
    win[0] = r0; win[1] = r1; win[2] = c0; win[3] = c1;
    // column pass
    for (int x = c0 + threadIdx.x; x < c1; x += MG_CTA) {
        int last = (empty && x == 0) ? -1 : -MG_INF; // empty + no window: scipy's phantom at (-1, 0)
        for (int y = r0; y < r1; y++) {
            if (L[y * W + x] == l) last = y;
            int d = y - last;
            G[y * W + x] = d > MG_INF ? MG_INF : d;
        }
        last = MG_INF;
        for (int y = r1 - 1; y >= r0; y--) {
            if (L[y * W + x] == l) last = y;
            int d = last - y;
            if (d < G[y * W + x]) G[y * W + x] = d;
        }
    }
    __syncthreads();
    // row pass
    int wh = r1 - r0, ww = c1 - c0;
    int mx = 0;
    for (int q = threadIdx.x; q < wh * ww; q += MG_CTA) {
        int yy = q / ww, y = r0 + yy, x = c0 + (q - yy * ww);
        const int32_t *g = G + y * W;
        int g0 = g[x];
        i64 best = g0 >= MG_INF ? ((i64)1 << 60) : (i64)g0 * g0;
        for (i64 k = 1; k * k < best; k++) {
            bool any = false;
            if (x - k >= c0) {
                any = true;
                int gv = g[x - k];
                if (gv < MG_INF) { i64 cnd = k * k + (i64)gv * gv; if (cnd < best) best = cnd; }
            }
            if (x + k < c1) {
                any = true;
                int gv = g[x + k];
                if (gv < MG_INF) { i64 cnd = k * k + (i64)gv * gv; if (cnd < best) best = cnd; }
            }
            if (!any) break;
        }
        int b = (int)best;
        D[y * W + x] = b;
        mx = max(mx, b);
    }
    *fill = mg_max_int(S, mx);
    return true;
}

__global__ void __launch_bounds__(MG_CTA) k_merge_labels(const int32_t *labels, int32_t *labels_out,
                                                         const maze_vignette_t *__restrict__ vig,
                                                         const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                         const int32_t *__restrict__ index,
                                                         const int32_t *__restrict__ index_off, int have_max,
                                                         double max_distance, double path_tolerance, int32_t *d2a,
                                                         int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch,
                                                         double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                                                         const int32_t *__restrict__ order)
{
    __shared__ MgShared S;
    __shared__ int s_scan[MG_CTA / 32 + 2];
    const int img = order ? order[blockIdx.x] : blockIdx.x; // longest vignettes first when an order is given
    maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, npx = H * W;
    const int32_t *L = labels + v.pix_off;
    int32_t *O = labels_out + v.pix_off;
    int32_t *A = d2a + v.pix_off, *B = d2b + v.pix_off, *G = gbuf + v.pix_off;
    int obj0 = lab_off[img];
    int bound = lab_off[img + 1] - obj0;
    if ((i64)obj0 + bound > n_obj_cap) bound = max(0, n_obj_cap - obj0);
    int32_t *idx = obj_scratch + 2 * (i64)obj0;
    uint32_t *mintab = (uint32_t *)(obj_scratch + 2 * (i64)obj0 + bound);
    if (threadIdx.x == 0) { n_merge[img] = 0; status[img] = MAZE_OK; }

    // merge_labels.py:55-57: index = sorted positive labels (or the caller's list)
    int n_idx = 0;
    if (index) {
        int i0 = index_off[img];
        n_idx = min(index_off[img + 1] - i0, bound);
        for (int j = threadIdx.x; j < n_idx; j += MG_CTA) idx[j] = index[i0 + j];
    } else {
        for (int j = threadIdx.x; j < bound; j += MG_CTA) mintab[j] = 0;
        __syncthreads();
        for (int p = threadIdx.x; p < npx; p += MG_CTA) {
            int l = L[p];
            if (l > 0 && l <= bound) mintab[l - 1] = 1;
        }
        __syncthreads();
        for (int j0 = 0; j0 < bound; j0 += MG_CTA) {
            int j = j0 + threadIdx.x;
            int f = (j < bound && mintab[j]) ? 1 : 0;
            // 512-thread exclusive scan
            int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            int inc = f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            if (threadIdx.x == 0) {
                int run = 0;
                for (int w = 0; w < MG_CTA / 32; w++) { int t = s_scan[w]; s_scan[w] = run; run += t; }
                s_scan[MG_CTA / 32] = run;
            }
            __syncthreads();
            if (f) idx[n_idx + s_scan[warp] + inc - 1] = j + 1;
            n_idx += s_scan[MG_CTA / 32];
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { index_state[2 * img] = n_idx; index_state[2 * img + 1] = 0; }
    if (n_idx < 2) return; // :59-60, nothing is written

    int pad = have_max ? (int)ceil(max_distance) + 1 : 0; // :70 and :20
    int l0 = idx[0];                                        // :66
    int head = 1;                                           // idx[head .. n_idx) is the remaining list
    for (int p = threadIdx.x; p < npx; p += MG_CTA)
        if (L[p] == l0) O[p] = l0;                          // :68
    __syncthreads();

    int win[4], fillA;
    if (threadIdx.x == 0) index_state[2 * img + 1] = 1;
    if (!mg_windowed_d2(S, L, H, W, l0, have_max, pad, G, A, win, &fillA)) {
        if (threadIdx.x == 0) status[img] = MAZE_ERR_TYPEERROR;
        return;
    }
    // :24 result = full(dist_sliced.max()); result[slices] = dist_sliced
    for (int p = threadIdx.x; p < npx; p += MG_CTA) {
        int y = p / W, x = p - y * W;
        if (y < win[0] || y >= win[1] || x < win[2] || x >= win[3]) A[p] = fillA;
    }
    const uint32_t maxd2 = (uint32_t)fillA; // :74 distmap.max()
    __syncthreads();

    int nm = 0;
    while (head < n_idx) { // :81
        // :83 per-label minimum of distmap, initial = max_dist
        for (int j = threadIdx.x; j < bound; j += MG_CTA) mintab[j] = 0xffffffffu;
        __syncthreads();
        for (int p0 = threadIdx.x; p0 < npx; p0 += 4 * MG_CTA) { // four independent pixels per thread in flight
            int l[4], a[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int p = p0 + u * MG_CTA;
                l[u] = p < npx ? L[p] : 0;
                a[u] = (l[u] > 0 && l[u] <= bound) ? A[p] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (l[u] > 0 && l[u] <= bound) atomicMin(mintab + (l[u] - 1), (uint32_t)a[u]);
        }
        __syncthreads();
        u64 best = ~0ull;
        for (int j = head + threadIdx.x; j < n_idx; j += MG_CTA) {
            int l = idx[j];
            uint32_t m = (l > 0 && l <= bound) ? mintab[l - 1] : 0xffffffffu;
            if (m > maxd2) m = maxd2;
            u64 key = ((u64)m << 32) | (uint32_t)j;
            best = key < best ? key : best;
        }
        best = mg_min_u64(S, best);
        int pos = (int)(best & 0xffffffffu);
        int cur_l = idx[pos]; // :84 index.pop(min_idx): order of the rest is preserved
        __syncthreads();
        for (int j0 = pos; j0 > head; j0 -= MG_CTA) {
            // shift idx[head .. pos) one to the right, from the top down
            int j = j0 - (int)threadIdx.x;
            int t = (j > head) ? idx[j - 1] : 0;
            __syncthreads();
            if (j > head) idx[j] = t;
            __syncthreads();
        }
        if (threadIdx.x == 0) idx[head] = cur_l; // popped entries stay in front, in pop order
        head++;
        if (threadIdx.x == 0) index_state[2 * img + 1] = head;
        __syncthreads();

        int winB[4], fillB;
        if (!mg_windowed_d2(S, L, H, W, cur_l, have_max, pad, G, B, winB, &fillB)) { // :87
            if (threadIdx.x == 0) { status[img] = MAZE_ERR_TYPEERROR; n_merge[img] = nm; }
            return;
        }
        const double sfillB = sqrt((double)fillB);
        double md = INFINITY;
        // :90-92 min over all pixels of sqrt(A) + sqrt(B).  sqrt(a) + sqrt(b) >= sqrt(a + b), so a pixel whose
        // a + b is clearly above the running minimum squared cannot lower it: most pixels need no sqrt at all
        const int sy = MG_CTA / W, sx = MG_CTA - sy * W;
        {
            int y = threadIdx.x / W, x = threadIdx.x - y * W;
            for (int p0 = threadIdx.x; p0 < npx; p0 += 4 * MG_CTA) {
                int a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * MG_CTA;
                    bool inb = !(y < winB[0] || y >= winB[1] || x < winB[2] || x >= winB[3]);
                    a[u] = p < npx ? A[p] : 0x3fffffff;
                    b[u] = (p < npx && inb) ? B[p] : fillB;
                    y += sy; x += sx;
                    if (x >= W) { x -= W; y++; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if ((double)a[u] + (double)b[u] > md * md * (1.0 + 1e-12)) continue;
                    double s = sqrt((double)a[u]) + sqrt((double)b[u]);
                    md = s < md ? s : md;
                }
            }
        }
        md = mg_min_double(S, md);
        if (have_max && md > max_distance) break; // :94-96
        const double lim = md + path_tolerance;
        if (threadIdx.x == 0 && merge_dist) merge_dist[obj0 + nm] = md; // :100
        nm++;
        {
            // :98, :103-106 (labelmap only ever holds l0) and :109-111.  sqrt(a+b) <= sqrt(a)+sqrt(b) <= sqrt(2(a+b))
            // decides most pixels on the integers; the float64 compare runs only in the narrow band between
            const double lim2 = lim * lim;
            int y = threadIdx.x / W, x = threadIdx.x - y * W;
            for (int p0 = threadIdx.x; p0 < npx; p0 += 4 * MG_CTA) {
                int a[4], b[4], l[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * MG_CTA;
                    bool inb = !(y < winB[0] || y >= winB[1] || x < winB[2] || x >= winB[3]);
                    a[u] = p < npx ? A[p] : 0;
                    b[u] = (p < npx && inb) ? B[p] : fillB;
                    l[u] = p < npx ? L[p] : 0;
                    y += sy; x += sx;
                    if (x >= W) { x -= W; y++; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * MG_CTA;
                    if (p >= npx) continue;
                    const double ab = (double)a[u] + (double)b[u];
                    bool fill = l[u] == cur_l;
                    if (!fill && !(ab > lim2 * (1.0 + 1e-12))) {
                        if (2.0 * ab < lim2 * (1.0 - 1e-12)) fill = true;
                        else fill = sqrt((double)a[u]) + sqrt((double)b[u]) <= lim;
                    }
                    if (fill) O[p] = l0;
                    if (b[u] < a[u]) A[p] = b[u];
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_merge[img] = nm;
}

extern "C" int maze_merge_labels(const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig, int n_img,
                                 const int32_t *lab_off, int n_obj_cap, const int32_t *index,
                                 const int32_t *index_off, int have_max, double max_distance, double path_tolerance,
                                 int32_t *d2a, int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch, double *merge_dist,
                                 int32_t *n_merge, int32_t *index_state, int32_t *status, const int32_t *order,
                                 void *stream)
{
    if (n_img <= 0) return MAZE_OK;
    if (index && !index_off) return MAZE_ERR_BADARG;
    MAZE_KERNEL(KID_MERGE_LABELS, (cudaStream_t)stream, k_merge_labels<<<n_img, MG_CTA, 0, (cudaStream_t)stream>>>(labels, labels_out, vig, lab_off, n_obj_cap, index,
                                                                index_off, have_max, max_distance, path_tolerance,
                                                                d2a, d2b, gbuf, obj_scratch, merge_dist, n_merge,
                                                                index_state, status, order));
    return MAZE_OK;
}
