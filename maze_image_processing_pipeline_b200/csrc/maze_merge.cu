// merge_labels on the device: one CTA -- or, for large vignettes, one thread-block CLUSTER of eight CTAs (cluster
// barrier between the steps, reductions through distributed shared memory) -- per vignette runs the reference's
// whole sequential loop.
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
#include <stdlib.h>

#include <cooperative_groups.h>

#include "maze_common.cuh"

namespace cg = cooperative_groups;

#define MG_CTA 1024
#define MG_INF (1 << 24)
#define MG_MAXCS 8
#define MG_SMIN 2048 /* labels per vignette whose per-iteration minima are taken in shared memory */

struct MgShared {
    int red_i[MG_CTA / 32][4];
    double red_d[MG_CTA / 32];
    u64 red_u[MG_CTA / 32];
    // one slot per CTA of the cluster (written through distributed shared memory into rank 0's copy)
    int cl_i[MG_MAXCS][4];
    double cl_d[MG_MAXCS];
    u64 cl_u[MG_MAXCS];
    int bbox[4];
    int ival;
    u64 uval;
    double dval;
};

// the CTAs of one vignette: CS = 1 is a plain CTA (barrier = __syncthreads), CS > 1 a thread-block cluster
template <int CS>
struct MgTeam {
    int rank, gt; // CTA rank in the cluster, thread index in the team
    static constexpr int GT = CS * MG_CTA;
    __device__ MgTeam()
    {
        rank = CS > 1 ? (int)cg::this_cluster().block_rank() : 0;
        gt = rank * MG_CTA + (int)threadIdx.x;
    }
    __device__ __forceinline__ void sync() const
    {
        if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
    }
    __device__ __forceinline__ MgShared *root(MgShared &S) const
    {
        return CS > 1 ? cg::this_cluster().map_shared_rank(&S, 0) : &S;
    }
};

// ---- team-wide reductions: warp -> CTA -> (cluster: slot in rank 0's shared memory, read back by every CTA) ----
template <int CS>
__device__ __forceinline__ void mg_bbox(const MgTeam<CS> &tm, MgShared &S, const int32_t *L, int H, int W, int l, int *out /*r0,r1,c0,c1 inclusive*/)
{
    int r0 = 0x7fffffff, r1 = -1, c0 = 0x7fffffff, c1 = -1;
    const int n = H * W;
    constexpr int GT = MgTeam<CS>::GT;
    {
        const int sy = GT / W, sx = GT - sy * W;
        int y = tm.gt / W, x = tm.gt - y * W;
        for (int p0 = tm.gt; p0 < n; p0 += 4 * GT) { // four independent loads in flight
            int v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (p0 + u * GT < n) ? L[p0 + u * GT] : l - 1;
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { S.red_i[warp][0] = r0; S.red_i[warp][1] = r1; S.red_i[warp][2] = c0; S.red_i[warp][3] = c1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) {
            r0 = min(r0, S.red_i[w][0]); r1 = max(r1, S.red_i[w][1]);
            c0 = min(c0, S.red_i[w][2]); c1 = max(c1, S.red_i[w][3]);
        }
        MgShared *R = tm.root(S);
        R->cl_i[tm.rank][0] = r0; R->cl_i[tm.rank][1] = r1; R->cl_i[tm.rank][2] = c0; R->cl_i[tm.rank][3] = c1;
    }
    tm.sync();
    if (threadIdx.x == 0) {
        const MgShared *R = tm.root(S);
        r0 = 0x7fffffff; r1 = -1; c0 = 0x7fffffff; c1 = -1;
        for (int k = 0; k < CS; k++) {
            r0 = min(r0, R->cl_i[k][0]); r1 = max(r1, R->cl_i[k][1]);
            c0 = min(c0, R->cl_i[k][2]); c1 = max(c1, R->cl_i[k][3]);
        }
        S.bbox[0] = r0; S.bbox[1] = r1; S.bbox[2] = c0; S.bbox[3] = c1;
    }
    tm.sync(); // (also: nobody overwrites rank 0's slots before everyone has read them)
    for (int j = 0; j < 4; j++) out[j] = S.bbox[j];
    __syncthreads();
}

template <int CS>
__device__ __forceinline__ int mg_max_int(const MgTeam<CS> &tm, MgShared &S, int v)
{
    v = __reduce_max_sync(FULL, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_i[warp][0] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = max(v, S.red_i[w][0]);
        tm.root(S)->cl_i[tm.rank][0] = v;
    }
    tm.sync();
    if (threadIdx.x == 0) {
        const MgShared *R = tm.root(S);
        v = R->cl_i[0][0];
        for (int k = 1; k < CS; k++) v = max(v, R->cl_i[k][0]);
        S.ival = v;
    }
    tm.sync();
    v = S.ival;
    __syncthreads();
    return v;
}

template <int CS>
__device__ __forceinline__ u64 mg_min_u64(const MgTeam<CS> &tm, MgShared &S, u64 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_u[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = S.red_u[w] < v ? S.red_u[w] : v;
        tm.root(S)->cl_u[tm.rank] = v;
    }
    tm.sync();
    if (threadIdx.x == 0) {
        const MgShared *R = tm.root(S);
        v = R->cl_u[0];
        for (int k = 1; k < CS; k++) v = R->cl_u[k] < v ? R->cl_u[k] : v;
        S.uval = v;
    }
    tm.sync();
    v = S.uval;
    __syncthreads();
    return v;
}

template <int CS>
__device__ __forceinline__ double mg_min_double(const MgTeam<CS> &tm, MgShared &S, double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_d[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MG_CTA / 32; w++) v = S.red_d[w] < v ? S.red_d[w] : v;
        tm.root(S)->cl_d[tm.rank] = v;
    }
    tm.sync();
    if (threadIdx.x == 0) {
        const MgShared *R = tm.root(S);
        v = R->cl_d[0];
        for (int k = 1; k < CS; k++) v = R->cl_d[k] < v ? R->cl_d[k] : v;
        S.dval = v;
    }
    tm.sync();
    v = S.dval;
    __syncthreads();
    return v;
}

// merge_labels.py:12-26.  Squared distances to the pixels of label l for the window
// (bbox enlarged by pad, clipped) into D; *fill receives the window maximum, win the window
// (half open).  Returns false where the reference raises TypeError (label absent, :19-20).
template <int CS>
__device__ bool mg_windowed_d2(const MgTeam<CS> &tm, MgShared &S, const int32_t *L, int H, int W, int l, int have_max, int pad,
                               int32_t *G, int32_t *D, int *win, int *fill)
{
    constexpr int GT = MgTeam<CS>::GT;
    int bb[4];
    mg_bbox(tm, S, L, H, W, l, bb);
    bool empty = bb[1] < 0;
    if (empty && have_max) return false;
    int r0 = 0, r1 = H, c0 = 0, c1 = W;
    if (have_max) {
        r0 = max(0, bb[0] - pad); r1 = (int)min((i64)H, (i64)bb[1] + 1 + pad);
        c0 = max(0, bb[2] - pad); c1 = (int)min((i64)W, (i64)bb[3] + 1 + pad);
    }
    win[0] = r0; win[1] = r1; win[2] = c0; win[3] = c1;
    // column pass
    for (int x = c0 + tm.gt; x < c1; x += GT) {
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
    tm.sync();
    // row pass
    int wh = r1 - r0, ww = c1 - c0;
    int mx = 0;
    for (int q = tm.gt; q < wh * ww; q += GT) {
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
    *fill = mg_max_int(tm, S, mx);
    return true;
}

template <int CS>
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
    __shared__ uint32_t s_min[MG_SMIN];
    const MgTeam<CS> tm;
    constexpr int GT = MgTeam<CS>::GT;
    const int gt = tm.gt;
    const int team = blockIdx.x / CS;
    const int img = order ? order[team] : team; // longest vignettes first when an order is given
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
    if (gt == 0) { n_merge[img] = 0; status[img] = MAZE_OK; }

    // merge_labels.py:55-57: index = sorted positive labels (or the caller's list)
    int n_idx = 0;
    if (index) {
        int i0 = index_off[img];
        n_idx = min(index_off[img + 1] - i0, bound);
        for (int j = gt; j < n_idx; j += GT) idx[j] = index[i0 + j];
    } else {
        for (int j = gt; j < bound; j += GT) mintab[j] = 0;
        tm.sync();
        for (int p = gt; p < npx; p += GT) {
            int l = L[p];
            if (l > 0 && l <= bound) mintab[l - 1] = 1;
        }
        tm.sync();
        if (tm.rank == 0) { // the list is short: the first CTA compacts it
            for (int j0 = 0; j0 < bound; j0 += MG_CTA) {
                int j = j0 + threadIdx.x;
                int f = (j < bound && mintab[j]) ? 1 : 0;
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
    }
    if (tm.rank == 0 && threadIdx.x == 0) { index_state[2 * img] = n_idx; index_state[2 * img + 1] = 0; }
    tm.sync();
    if (CS > 1) n_idx = *(volatile int32_t *)(index_state + 2 * img); // (the other CTAs did not build the list)
    if (n_idx < 2) return; // :59-60, nothing is written

    int pad = have_max ? (int)ceil(max_distance) + 1 : 0; // :70 and :20
    int l0 = idx[0];                                        // :66
    int head = 1;                                           // idx[head .. n_idx) is the remaining list
    for (int p = gt; p < npx; p += GT)
        if (L[p] == l0) O[p] = l0;                          // :68
    tm.sync();

    int win[4], fillA;
    if (gt == 0) index_state[2 * img + 1] = 1;
    if (!mg_windowed_d2(tm, S, L, H, W, l0, have_max, pad, G, A, win, &fillA)) {
        if (gt == 0) status[img] = MAZE_ERR_TYPEERROR;
        return;
    }
    // :24 result = full(dist_sliced.max()); result[slices] = dist_sliced
    for (int p = gt; p < npx; p += GT) {
        int y = p / W, x = p - y * W;
        if (y < win[0] || y >= win[1] || x < win[2] || x >= win[3]) A[p] = fillA;
    }
    const uint32_t maxd2 = (uint32_t)fillA; // :74 distmap.max()
    tm.sync();

    int nm = 0;
    while (head < n_idx) { // :81
        // :83 per-label minimum of distmap, initial = max_dist.  The minima are taken in SHARED memory first (every
        // labelled pixel of the vignette is one atomicMin on one of a handful of addresses: in global memory they
        // serialise at the L2) and reach the global table with one atomic per label and CTA
        const bool smin = bound <= MG_SMIN;
        if (smin)
            for (int j = threadIdx.x; j < bound; j += MG_CTA) s_min[j] = 0xffffffffu;
        for (int j = gt; j < bound; j += GT) mintab[j] = 0xffffffffu;
        tm.sync();
        for (int p0 = gt; p0 < npx; p0 += 4 * GT) { // four independent pixels per thread in flight
            int l[4], a[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int p = p0 + u * GT;
                l[u] = p < npx ? L[p] : 0;
                a[u] = (l[u] > 0 && l[u] <= bound) ? A[p] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (l[u] > 0 && l[u] <= bound) atomicMin((smin ? s_min : mintab) + (l[u] - 1), (uint32_t)a[u]);
        }
        if (smin) {
            __syncthreads();
            for (int j = threadIdx.x; j < bound; j += MG_CTA)
                if (s_min[j] != 0xffffffffu) atomicMin(mintab + j, s_min[j]);
        }
        tm.sync();
        u64 best = ~0ull;
        for (int j = head + gt; j < n_idx; j += GT) {
            int l = idx[j];
            uint32_t m = (l > 0 && l <= bound) ? mintab[l - 1] : 0xffffffffu;
            if (m > maxd2) m = maxd2;
            u64 key = ((u64)m << 32) | (uint32_t)j;
            best = key < best ? key : best;
        }
        best = mg_min_u64(tm, S, best);
        int pos = (int)(best & 0xffffffffu);
        int cur_l = idx[pos]; // :84 index.pop(min_idx): order of the rest is preserved
        tm.sync();
        if (tm.rank == 0) {
            for (int j0 = pos; j0 > head; j0 -= MG_CTA) {
                // shift idx[head .. pos) one to the right, from the top down
                int j = j0 - (int)threadIdx.x;
                int t = (j > head) ? idx[j - 1] : 0;
                __syncthreads();
                if (j > head) idx[j] = t;
                __syncthreads();
            }
            if (threadIdx.x == 0) idx[head] = cur_l; // popped entries stay in front, in pop order
        }
        head++;
        if (gt == 0) index_state[2 * img + 1] = head;
        tm.sync();

        int winB[4], fillB;
        if (!mg_windowed_d2(tm, S, L, H, W, cur_l, have_max, pad, G, B, winB, &fillB)) { // :87
            if (gt == 0) { status[img] = MAZE_ERR_TYPEERROR; n_merge[img] = nm; }
            return;
        }
        double md = INFINITY;
        // :90-92 min over all pixels of sqrt(A) + sqrt(B).  sqrt(a) + sqrt(b) >= sqrt(a + b), so a pixel whose
        // a + b is clearly above the running minimum squared cannot lower it: most pixels need no sqrt at all
        const int sy = GT / W, sx = GT - sy * W;
        {
            int y = gt / W, x = gt - y * W;
            for (int p0 = gt; p0 < npx; p0 += 4 * GT) {
                int a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * GT;
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
        md = mg_min_double(tm, S, md);
        if (have_max && md > max_distance) break; // :94-96
        const double lim = md + path_tolerance;
        if (gt == 0 && merge_dist) merge_dist[obj0 + nm] = md; // :100
        nm++;
        {
            // :98, :103-106 (labelmap only ever holds l0) and :109-111.  sqrt(a+b) <= sqrt(a)+sqrt(b) <= sqrt(2(a+b))
            // decides most pixels on the integers; the float64 compare runs only in the narrow band between
            const double lim2 = lim * lim;
            int y = gt / W, x = gt - y * W;
            for (int p0 = gt; p0 < npx; p0 += 4 * GT) {
                int a[4], b[4], l[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * GT;
                    bool inb = !(y < winB[0] || y >= winB[1] || x < winB[2] || x >= winB[3]);
                    a[u] = p < npx ? A[p] : 0;
                    b[u] = (p < npx && inb) ? B[p] : fillB;
                    l[u] = p < npx ? L[p] : 0;
                    y += sy; x += sx;
                    if (x >= W) { x -= W; y++; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int p = p0 + u * GT;
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
        tm.sync();
    }
    if (gt == 0) n_merge[img] = nm;
}

template <int CS>
static int launch_merge(int n_teams, cudaStream_t s, const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig,
                        const int32_t *lab_off, int n_obj_cap, const int32_t *index, const int32_t *index_off, int have_max,
                        double max_distance, double path_tolerance, int32_t *d2a, int32_t *d2b, int32_t *gbuf,
                        int32_t *obj_scratch, double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                        const int32_t *order)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_teams * CS, 1, 1);
    cfg.blockDim = dim3(MG_CTA, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    maze_prof_begin(KID_MERGE_LABELS, s);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_merge_labels<CS>, labels, labels_out, vig, lab_off, n_obj_cap, index, index_off,
                                       have_max, max_distance, path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist,
                                       n_merge, index_state, status, order);
    maze_prof_end(KID_MERGE_LABELS, s);
    if (e != cudaSuccess) {
        maze_set_err(e, "KID_MERGE_LABELS");
        return MAZE_ERR_CUDA;
    }
    MAZE_LAUNCH_CHECK("KID_MERGE_LABELS");
    return MAZE_OK;
}

int maze_merge_windowed_launch(int n_teams, cudaStream_t s, const int32_t *labels, int32_t *labels_out,
                               const maze_vignette_t *vig, const int32_t *lab_off, int n_obj_cap, double max_distance,
                               double path_tolerance, int32_t *d2a, int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch,
                               double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                               const int32_t *order); // maze_merge_win.cu

// n_order vignettes (order[0 .. n_order), or all n_img in index order when order is NULL), each worked on by
// cluster_size CTAs (1, or 8: a thread-block cluster with DSMEM reductions -- for large vignettes)
extern "C" int maze_merge_labels_ex(const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig, int n_img,
                                    const int32_t *lab_off, int n_obj_cap, const int32_t *index,
                                    const int32_t *index_off, int have_max, double max_distance, double path_tolerance,
                                    int32_t *d2a, int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch, double *merge_dist,
                                    int32_t *n_merge, int32_t *index_state, int32_t *status, const int32_t *order,
                                    int n_order, int cluster_size, void *stream)
{
    if (n_img <= 0) return MAZE_OK;
    if (index && !index_off) return MAZE_ERR_BADARG;
    if (cluster_size != 1 && cluster_size != MG_MAXCS) return MAZE_ERR_BADARG;
    const int n_teams = order ? n_order : n_img;
    if (n_teams <= 0) return MAZE_OK;
    // "sorted positive labels" with a maximum distance (the pipeline's call): the windowed kernel
    if (!index && have_max) {
        const char *e = getenv("MAZE_MERGE_WINDOWED");
        if (!e || e[0] != '0')
            return maze_merge_windowed_launch(n_teams, (cudaStream_t)stream, labels, labels_out, vig, lab_off, n_obj_cap,
                                              max_distance, path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist, n_merge,
                                              index_state, status, order);
    }
    if (cluster_size == 1)
        return launch_merge<1>(n_teams, (cudaStream_t)stream, labels, labels_out, vig, lab_off, n_obj_cap, index, index_off,
                               have_max, max_distance, path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist, n_merge,
                               index_state, status, order);
    return launch_merge<MG_MAXCS>(n_teams, (cudaStream_t)stream, labels, labels_out, vig, lab_off, n_obj_cap, index, index_off,
                                  have_max, max_distance, path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist, n_merge,
                                  index_state, status, order);
}

extern "C" int maze_merge_labels(const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig, int n_img,
                                 const int32_t *lab_off, int n_obj_cap, const int32_t *index,
                                 const int32_t *index_off, int have_max, double max_distance, double path_tolerance,
                                 int32_t *d2a, int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch, double *merge_dist,
                                 int32_t *n_merge, int32_t *index_state, int32_t *status, const int32_t *order,
                                 void *stream)
{
    return maze_merge_labels_ex(labels, labels_out, vig, n_img, lab_off, n_obj_cap, index, index_off, have_max, max_distance,
                                path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist, n_merge, index_state, status, order,
                                n_img, 1, stream);
}
