// merge_labels with a maximum distance, WINDOWED: one CTA per vignette runs the reference's sequential loop, but every
// iteration only touches the distance window of the label it pops (bounding box + ceil(max_distance) + 1), not the
// whole vignette.
//
// Reference behaviour restated: maze_ipp/merge_labels.py:29-113 (helpers :7-26) with index = None and a max_distance
// (what maze_ipp/loki/pipeline.py:451-457 calls, aliased).  Same numbers as k_merge_labels (maze_merge.cu: exact
// integer squared distances, float64 sqrt / add / compare); what changes is WHERE they are evaluated:
//
//  * `cur_distmap` is `dist_sliced.max()` (= fillB) outside the window (:24-25), so
//      - `cur_distmap < distmap` outside the window lowers distmap to fillB wherever it is larger: a scalar CAP that is
//        applied when distmap is read (distmap = min(stored, cap)), never written;
//      - `sum_distmap.min()` outside the window is sqrt(0) + sqrt(fillB) iff a pixel with distmap == 0 lies outside
//        the window -- the pixels of the labels merged so far, whose bounding box is tracked -- and otherwise some
//        pixel INSIDE the window has distmap == 0 and gives a sum <= sqrt(fillB) <= every outside sum;
//      - `sum_distmap <= merge_dist + path_tolerance` outside the window needs sqrt(fillB) <= merge_dist +
//        path_tolerance: rare (the window maximum is at least pad as soon as one margin is unclipped), and then a pass
//        over the rest of the vignette evaluates it exactly;
//  * the loop ends without a distance map when the popped label's minimum of distmap exceeds max_distance^2, and a
//    vignette in which no foreign pixel comes within max_distance of l0 ends before the first map (proofs at the two tests in the kernel);
//  * `distmap[labels == l].min(initial=max_dist)` is kept per label and lowered by the pixels whose distmap an
//    iteration lowers; in the aliased call a label can LOSE pixels to a bridge: if the lost pixel sat on the label's
//    bounding box or held its minimum, the label is recomputed over its old bounding box before the next pop.
#include <math.h>

#include "maze_common.cuh"

#ifndef MW_CTA
#define MW_CTA 1024 /* one CTA per SM: the long vignettes decide the makespan, they get the whole SM */
#endif
#define MW_MINB (1024 / MW_CTA)
#define MW_NW (MW_CTA / 32)
#define MW_INF (1 << 24)
#define MW_WCAP 6144 /* window pixels whose column distances and squared distances stay in shared memory */
#define MW_TCAP 1024 /* labels per vignette whose tables (box, minimum, flag) stay in shared memory */
#define MW_DIRTY 32
#define MW_NEAR_BUDGET 2048 /* pixels of other labels whose neighbourhood is searched for l0 before its distance map is built */
#define MW_PARTS 8            /* CTAs of the preparation kernel per vignette */
#define MW_PREP_CTA 256
#define MW_BIG 0x7f7f7f7f     /* "not computed yet": every stored distance is below it */
#define MW_FLIP 0x3fffffff    /* the preparation kernel keeps box minima as MW_FLIP - v: every update is an atomicMax on -1 */
#define MW_SMEM_BYTES ((2 * MW_WCAP + 6 * MW_TCAP) * 4)

#ifdef MW_TIMING
// per-vignette phase times (ns, thread 0): total, tables, edt0, mins0, dirty, pop, edt, summin, fill, outside | iterations
__device__ long long mw_dbg[8192 * 16];
__shared__ int mw_img; // (vignette of this CTA, for the laps inside mw_edt)
__device__ __forceinline__ long long mw_now()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MW_T0() long long t__ = mw_now(), tstart__ = t__; if (threadIdx.x == 0) mw_img = img; __syncthreads()
#define MW_ELAP(k) do { __syncthreads(); long long n__ = mw_now(); if (threadIdx.x == 0) mw_dbg[(i64)mw_img * 16 + (k)] += n__ - te__; te__ = n__; } while (0)
#define MW_ET0() long long te__ = mw_now()
#define MW_LAP(k) do { long long n__ = mw_now(); if (threadIdx.x == 0) mw_dbg[(i64)img * 16 + (k)] += n__ - t__; t__ = n__; } while (0)
#define MW_END(it) do { if (threadIdx.x == 0) { mw_dbg[(i64)img * 16] = mw_now() - tstart__; mw_dbg[(i64)img * 16 + 10] = (it); \
                                               mw_dbg[(i64)img * 16 + 11] = blockIdx.x; } } while (0)
extern "C" int maze_merge_debug_read(long long *host, int n_img)
{
    return cudaMemcpyFromSymbol(host, mw_dbg, sizeof(long long) * 16 * (size_t)n_img) == cudaSuccess ? 0 : -1;
}
extern "C" int maze_merge_debug_clear(void)
{
    static long long z[8192 * 16];
    return cudaMemcpyToSymbol(mw_dbg, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#else
#define MW_T0()
#define MW_LAP(k)
#define MW_END(it)
#define MW_ELAP(k)
#define MW_ET0()
#endif

struct MwShared {
    u64 red_u[MW_NW];
    double red_d[MW_NW];
    int red_i[MW_NW];
    u64 uval;
    double dval;
    int ival;
    int n_dirty;
    int dirty[MW_DIRTY];
    int near, tested;
    int scan[MW_NW + 1];
};

__device__ __forceinline__ int mw_max_int(MwShared &S, int v)
{
    v = __reduce_max_sync(FULL, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_i[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int t = lane < MW_NW ? S.red_i[lane] : (int)0x80000000;
        t = __reduce_max_sync(FULL, t);
        if (lane == 0) S.ival = t;
    }
    __syncthreads();
    return S.ival;
}

__device__ __forceinline__ u64 mw_min_u64(MwShared &S, u64 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_u[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u64 t = lane < MW_NW ? S.red_u[lane] : ~0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            u64 o = __shfl_xor_sync(FULL, t, d);
            t = o < t ? o : t;
        }
        if (lane == 0) S.uval = t;
    }
    __syncthreads();
    return S.uval;
}

__device__ __forceinline__ double mw_min_double(MwShared &S, double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double o = __shfl_xor_sync(FULL, v, d);
        v = o < v ? o : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) S.red_d[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < MW_NW ? S.red_d[lane] : INFINITY;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double o = __shfl_xor_sync(FULL, t, d);
            t = o < t ? o : t;
        }
        if (lane == 0) S.dval = t;
    }
    __syncthreads();
    return S.dval;
}

// (Rebuild of the tables when one bridge took pixels from more labels than the kernel lists; the first boxes come
// from k_mw_prepare.)  Bounding boxes (r0, r1, c0, c1 inclusive) of every label 1 .. bound -- and, WITH_MIN, the
// minimum of A over its pixels -- in one pass over the vignette: a warp per row, four 32-pixel groups in flight; the lanes of a group that
// hold the same label elect a leader, which knows the group's column extent from the match mask.
template <bool WITH_MIN>
__device__ void mw_scan_all(const int32_t *L, const int32_t *A, int H, int W, int bound, int *box, uint32_t *mintab)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int y = warp; y < H; y += MW_NW) {
        const int32_t *row = L + (i64)y * W;
        for (int xb = 0; xb < W; xb += 128) {
            int l[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int x = xb + 32 * u + lane;
                l[u] = x < W ? row[x] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int x0 = xb + 32 * u;
                const bool ok = l[u] > 0 && l[u] <= bound;
                if (!__ballot_sync(FULL, ok)) continue;
                const unsigned grp = __match_any_sync(FULL, ok ? l[u] : 0);
                if (ok) {
                    const int lead = __ffs(grp) - 1;
                    if (WITH_MIN) {
                        uint32_t a = (uint32_t)A[(i64)y * W + x0 + lane];
                        a = __reduce_min_sync(grp, a);
                        if (lane == lead) atomicMin(mintab + (l[u] - 1), a);
                    }
                    if (lane == lead) {
                        int *b = box + 4 * (l[u] - 1);
                        atomicMin(b + 0, y);
                        atomicMax(b + 1, y);
                        atomicMin(b + 2, x0 + lead);
                        atomicMax(b + 3, x0 + 31 - __clz(grp));
                    }
                }
            }
        }
    }
}

// merge_labels.py:12-26 for a label with a known bounding box: squared distances to the pixels of label l inside the
// window win (r0, r1, c0, c1; half open) into dst[(y - r0) * ds + (x - c0)]; Gp (stride gs) takes the column
// distances.  Returns the window maximum (`dist_sliced.max()`).  The label has at least one pixel in the window.
__device__ int mw_edt(MwShared &S, const int32_t *L, int W, int l, const int *win, int32_t *Gp, int gs, int32_t *dst, int ds)
{
    const int r0 = win[0], c0 = win[2];
    const int wh = win[1] - r0, ww = win[3] - c0, npw = wh * ww;
    if (npw > MW_WCAP && wh <= 32767 && ww <= 32767) {
        // LARGE window.  Horizontal distances first (a warp per row, coalesced), then the lower envelope of the
        // parabolas (y - v)^2 + g[v]^2 down every COLUMN, one thread per column: the lanes of a warp walk neighbouring
        // columns in step, so the loads of g, the stack of the envelope and the distances written at the end are all
        // row-major accesses of neighbouring words.  Exact in integers.
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        MW_ET0();
        for (int yy = warp; yy < wh; yy += MW_NW) {
            const int32_t *src = L + (i64)(r0 + yy) * W + c0;
            int32_t *g = Gp + (i64)yy * gs;
            int last = -MW_INF; // column of the last pixel of the label seen so far
            for (int xb = 0; xb < ww; xb += 256) {
                bool is[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int x = xb + 32 * u + lane;
                    is[u] = x < ww && src[x] == l;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int x0 = xb + 32 * u, x = x0 + lane;
                    if (x0 >= ww) break;
                    const unsigned bits = __ballot_sync(FULL, is[u]);
                    const unsigned lo = bits & (0xffffffffu >> (31 - lane)); // the label's pixels at or left of this lane
                    const int d = lo ? lane - (31 - __clz(lo)) : min(x - last, MW_INF);
                    if (x < ww) g[x] = d;
                    if (bits) last = x0 + 31 - __clz(bits);
                }
            }
            int nxt = 2 * MW_INF; // column of the nearest pixel of the label to the right of the group
            for (int xb = ((ww - 1) >> 8) << 8; xb >= 0; xb -= 256) {
                int dl[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int x = xb + 32 * u + lane;
                    dl[u] = x < ww ? g[x] : MW_INF;
                }
#pragma unroll
                for (int u = 7; u >= 0; u--) {
                    const int x0 = xb + 32 * u, x = x0 + lane;
                    if (x0 >= ww) continue;
                    const unsigned bits = __ballot_sync(FULL, dl[u] == 0);
                    const unsigned hi = bits & (0xffffffffu << lane);
                    const int d = hi ? __ffs(hi) - 1 - lane : min(nxt - x, MW_INF);
                    if (x < ww && d < dl[u]) g[x] = d;
                    if (bits) nxt = x0 + __ffs(bits) - 1;
                }
            }
        }
        __syncthreads();
        MW_ELAP(12);
        // The stack of the envelope lives in the column itself: entry k = v | start << 16 in dst[k][x] (start = first
        // row where the parabola at v is strictly below its predecessor, clamped to the window) and g[v] in Gp[k][x]
        // (k <= the row being read).  The second sweep runs bottom up, so that the distances it writes over dst[q][x]
        // only hit stack entries that are no longer needed (k <= start[k] <= q).
        // (sides <= 32767: every intermediate below fits 32 bits -- g^2 + q^2 < 2^31, st * den < 32767 * 65534 < 2^31)
        int mx = 0;
        for (int xx = threadIdx.x; xx < ww; xx += MW_CTA) {
            int32_t *gc = Gp + xx;
            int32_t *col = dst + xx;
            int k = -1, vt = 0, st = 0, gt = 0;
            int ft = 0; // g[vt]^2 + vt^2
            int gn8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) gn8[u] = u < wh ? gc[u * gs] : MW_INF;
            for (int q0 = 0; q0 < wh; q0 += 8) {
                int g8[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    g8[u] = gn8[u];
                    gn8[u] = q0 + 8 + u < wh ? gc[(q0 + 8 + u) * gs] : MW_INF; // (the stack writes stay at rows <= q0 + 7)
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int q = q0 + u, gq = g8[u];
                    if (gq >= MW_INF) continue;
                    const int fq = gq * gq + q * q;
                    int s = 0;
                    while (k >= 0) {
                        // the parabola at q is strictly below the one at vt from floor(num / den) + 1 on; it leaves the
                        // top of the stack in place iff that is beyond the top's start: floor(num / den) >= st
                        const int num = fq - ft, den = 2 * (q - vt);
                        if (num >= st * den) {
                            s = min((int)((unsigned)num / (unsigned)den) + 1, wh); // (num >= 0 here)
                            break;
                        }
                        if (--k >= 0) {
                            const uint32_t e = (uint32_t)col[k * ds];
                            vt = (int)(e & 0xffffu); st = (int)(e >> 16); gt = gc[k * gs];
                            ft = gt * gt + vt * vt;
                        }
                    }
                    if (k < 0) s = 0;
                    else if (s >= wh) continue; // never the minimum inside the window
                    k++;
                    col[k * ds] = (int32_t)((uint32_t)q | ((uint32_t)s << 16));
                    gc[k * gs] = gq;
                    vt = q; st = s; gt = gq; ft = fq;
                }
            }
            // (the entry below the current one is loaded ahead of its use)
            int nk = k - 1;
            uint32_t en = nk >= 0 ? (uint32_t)col[nk * ds] : 0u;
            int gn = nk >= 0 ? gc[nk * gs] : 0;
            for (int q = wh - 1; q >= 0; q--) {
                while (st > q) {
                    k = nk; vt = (int)(en & 0xffffu); st = (int)(en >> 16); gt = gn;
                    nk = k - 1;
                    if (nk >= 0) { en = (uint32_t)col[nk * ds]; gn = gc[nk * gs]; }
                }
                const int d = (q - vt) * (q - vt) + gt * gt;
                col[q * ds] = d;
                mx = max(mx, d);
            }
        }
        MW_ELAP(13);
        return mw_max_int(S, mx);
    }
    const int nseg = (wh + 31) >> 5;
    const int items = nseg * ww;
    // columns in segments of 32 rows: the label's rows of a (segment, column) as one word -> row s of dst (free until
    // the row pass; nseg <= wh)
    for (int t = threadIdx.x; t < items; t += MW_CTA) {
        const int s = t / ww, xx = t - s * ww;
        const int y0 = s << 5;
        const int32_t *src = L + (i64)(r0 + y0) * W + c0 + xx;
        unsigned bits = 0;
        if (y0 + 32 <= wh) {
#pragma unroll
            for (int r = 0; r < 32; r++) bits |= (src[(i64)r * W] == l ? 1u : 0u) << r;
        } else {
            for (int r = 0; r < wh - y0; r++) bits |= (src[(i64)r * W] == l ? 1u : 0u) << r;
        }
        dst[(i64)s * ds + xx] = (int)bits;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < items; t += MW_CTA) {
        const int s = t / ww, xx = t - s * ww;
        const int y0 = s << 5, y1 = min(wh, y0 + 32);
        const unsigned bits = (unsigned)dst[(i64)s * ds + xx];
        int above = -MW_INF, below = MW_INF; // nearest row of the label before / after this segment
        for (int k = s - 1; k >= 0; k--) {
            const unsigned b = (unsigned)dst[(i64)k * ds + xx];
            if (b) { above = (k << 5) + 31 - __clz(b); break; }
        }
        for (int k = s + 1; k < nseg; k++) {
            const unsigned b = (unsigned)dst[(i64)k * ds + xx];
            if (b) { below = (k << 5) + __ffs(b) - 1; break; }
        }
        for (int yy = y0; yy < y1; yy++) {
            const int r = yy - y0;
            int d = 0;
            if (!((bits >> r) & 1)) {
                const unsigned lo = bits & ((1u << r) - 1u), hi = (bits >> r) >> 1;
                const int up = lo ? r - (31 - __clz(lo)) : yy - above;
                const int dn = hi ? __ffs(hi) : below - yy;
                d = min(min(up, dn), MW_INF);
            }
            Gp[(i64)yy * gs + xx] = d;
        }
    }
    __syncthreads();
    int mx = 0;
    // row pass: min over the row of k^2 + g^2 (integers); four pixels per thread with their first loads in flight
    for (int q0 = threadIdx.x; q0 < npw; q0 += 4 * MW_CTA) {
        int g0[4], yy[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int q = q0 + u * MW_CTA;
            yy[u] = q / ww; xx[u] = q - yy[u] * ww;
            g0[u] = q < npw ? Gp[(i64)yy[u] * gs + xx[u]] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (q0 + u * MW_CTA >= npw) break;
            const int32_t *g = Gp + (i64)yy[u] * gs;
            const int x = xx[u];
            i64 best = g0[u] >= MW_INF ? ((i64)1 << 60) : (i64)g0[u] * g0[u];
            for (i64 k = 1; k * k < best; k++) {
                bool any = false;
                if (x - k >= 0) {
                    any = true;
                    int gv = g[x - k];
                    if (gv < MW_INF) { i64 cnd = k * k + (i64)gv * gv; if (cnd < best) best = cnd; }
                }
                if (x + k < ww) {
                    any = true;
                    int gv = g[x + k];
                    if (gv < MW_INF) { i64 cnd = k * k + (i64)gv * gv; if (cnd < best) best = cnd; }
                }
                if (!any) break;
            }
            const int b = (int)best;
            // (rows 0 .. nseg of dst held the segment words until every thread passed the barrier above)
            dst[(i64)yy[u] * ds + x] = b;
            mx = max(mx, b);
        }
    }
    return mw_max_int(S, mx);
}

__device__ __forceinline__ void mw_mark_dirty(MwShared &S, int *dflag, int l)
{
    if (dflag[l - 1] == 0 && atomicExch(dflag + (l - 1), 1) == 0) {
        int k = atomicAdd(&S.n_dirty, 1);
        if (k < MW_DIRTY) S.dirty[k] = l;
    }
}

// a pixel of label l (not l0, not the popped label) is about to be overwritten by a bridge (aliased call): the label's
// box / minimum must be recomputed if the pixel sits on the box or holds the minimum
__device__ __forceinline__ void mw_lose_pixel(MwShared &S, const int *box, const uint32_t *mintab, int *dflag, int l, int y, int x,
                                              uint32_t a_raw)
{
    const int *b = box + 4 * (l - 1);
    if (y == b[0] || y == b[1] || x == b[2] || x == b[3] || a_raw <= mintab[l - 1]) mw_mark_dirty(S, dflag, l);
}

// Before the loop, with the whole GPU instead of one CTA per vignette: A = MW_BIG everywhere, and the bounding box of
// every label (obj_scratch + 2 n_obj_cap + 4 (lab_off[i] + l - 1): MW_FLIP - r0, r1, MW_FLIP - c0, c1, all -1 = absent;
// the caller clears the table to 0xff).  MW_PARTS CTAs per vignette, a slab of rows each; a warp walks rows and keeps
// the box of the label it saw last in registers, so that a large object costs a handful of atomics per warp.
__global__ void __launch_bounds__(MW_PREP_CTA) k_mw_prepare(const int32_t *__restrict__ labels,
                                                            const maze_vignette_t *__restrict__ vig,
                                                            const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                            int32_t *__restrict__ d2a, int32_t *obj_scratch,
                                                            const int32_t *__restrict__ order)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = order ? order[blockIdx.x] : blockIdx.x;
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w;
    const int obj0 = lab_off[img];
    int bound = lab_off[img + 1] - obj0;
    if ((i64)obj0 + bound > n_obj_cap) bound = max(0, n_obj_cap - obj0);
    const int rp = (H + MW_PARTS - 1) / MW_PARTS;
    const int y0 = blockIdx.y * rp, y1 = min(H, y0 + rp);
    if (y0 >= y1) return;
    const int32_t *L = labels + v.pix_off;
    int32_t *A = d2a + v.pix_off;
    int *gbox = obj_scratch + 2 * (i64)n_obj_cap + 4 * (i64)obj0;
    if (bound >= 2) // (fewer than two labels: the loop returns before it reads any distance, :59-60)
        for (i64 p = (i64)y0 * W + threadIdx.x, p1 = (i64)y1 * W; p < p1; p += MW_PREP_CTA) A[p] = MW_BIG;

    int cl = 0, r0 = 0, r1 = 0, c0 = 0, c1 = 0; // (warp-uniform)
    auto flush = [&]() {
        if (cl > 0 && lane == 0) {
            int *b = gbox + 4 * (cl - 1);
            atomicMax(b + 0, MW_FLIP - r0);
            atomicMax(b + 1, r1);
            atomicMax(b + 2, MW_FLIP - c0);
            atomicMax(b + 3, c1);
        }
    };
    for (int y = y0 + warp; y < y1; y += MW_PREP_CTA / 32) {
        const int32_t *row = L + (i64)y * W;
        for (int xb = 0; xb < W; xb += 256) {
            int l[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int x = xb + 32 * u + lane;
                l[u] = x < W ? row[x] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int x0 = xb + 32 * u;
                unsigned todo = __ballot_sync(FULL, l[u] > 0 && l[u] <= bound);
                while (todo) {
                    const int lv = __shfl_sync(FULL, l[u], __ffs(todo) - 1);
                    const unsigned m = __ballot_sync(FULL, l[u] == lv);
                    todo &= ~m;
                    const int xa = x0 + __ffs(m) - 1, xz = x0 + 31 - __clz(m);
                    if (lv != cl) {
                        flush();
                        cl = lv; r0 = r1 = y; c0 = xa; c1 = xz;
                    } else {
                        r0 = min(r0, y); r1 = max(r1, y); c0 = min(c0, xa); c1 = max(c1, xz);
                    }
                }
            }
        }
    }
    flush();
}

__global__ void __launch_bounds__(MW_CTA, MW_MINB) k_merge_windowed(const int32_t *labels, int32_t *labels_out,
                                                              const maze_vignette_t *__restrict__ vig,
                                                              const int32_t *__restrict__ lab_off, int n_obj_cap,
                                                              double max_distance, double path_tolerance, int32_t *d2a,
                                                              int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch,
                                                              double *merge_dist, int32_t *n_merge, int32_t *index_state,
                                                              int32_t *status, const int32_t *__restrict__ order)
{
    extern __shared__ __align__(16) int32_t mw_smem[];
    __shared__ MwShared S;
    int32_t *sB = mw_smem, *sG = mw_smem + MW_WCAP, *sT = mw_smem + 2 * MW_WCAP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = order ? order[blockIdx.x] : blockIdx.x;
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w;
    const int32_t *L = labels + v.pix_off;
    int32_t *O = labels_out + v.pix_off;
    int32_t *A = d2a + v.pix_off, *B = d2b + v.pix_off, *G = gbuf + v.pix_off;
    const bool aliased = labels == labels_out;
    const int obj0 = lab_off[img];
    int bound = lab_off[img + 1] - obj0;
    if ((i64)obj0 + bound > n_obj_cap) bound = max(0, n_obj_cap - obj0);
    int32_t *idx = obj_scratch + 2 * (i64)obj0;
    const bool stab = bound <= MW_TCAP;
    int *box = stab ? sT : obj_scratch + 2 * (i64)n_obj_cap + 4 * (i64)obj0;
    uint32_t *mintab = stab ? (uint32_t *)(sT + 4 * MW_TCAP) : (uint32_t *)(idx + bound);
    int *dflag = stab ? sT + 5 * MW_TCAP : obj_scratch + 6 * (i64)n_obj_cap + obj0;
    if (tid == 0) { n_merge[img] = 0; status[img] = MAZE_OK; S.n_dirty = 0; }
    MW_T0();

    {
        // label boxes of k_mw_prepare (atomicMax form) -> r0, r1, c0, c1
        const int *gbox = obj_scratch + 2 * (i64)n_obj_cap + 4 * (i64)obj0;
        for (int j = tid; j < bound; j += MW_CTA) {
            const int t0 = gbox[4 * j], t1 = gbox[4 * j + 1], t2 = gbox[4 * j + 2], t3 = gbox[4 * j + 3];
            const bool none = t1 < 0;
            box[4 * j] = none ? 0x7fffffff : MW_FLIP - t0; box[4 * j + 1] = none ? -1 : t1;
            box[4 * j + 2] = none ? 0x7fffffff : MW_FLIP - t2; box[4 * j + 3] = none ? -1 : t3;
            mintab[j] = 0xffffffffu;
            dflag[j] = 0;
        }
    }
    __syncthreads();
    // merge_labels.py:55-57: index = sorted positive labels
    int n_idx = 0;
    for (int j0 = 0; j0 < bound; j0 += MW_CTA) {
        const int j = j0 + tid;
        const int f = (j < bound && box[4 * j + 1] >= 0) ? 1 : 0;
        int inc = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) S.scan[warp] = inc;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < MW_NW; w++) { int t = S.scan[w]; S.scan[w] = run; run += t; }
            S.scan[MW_NW] = run;
        }
        __syncthreads();
        if (f) idx[n_idx + S.scan[warp] + inc - 1] = j + 1;
        n_idx += S.scan[MW_NW];
        __syncthreads();
    }
    if (tid == 0) { index_state[2 * img] = n_idx; index_state[2 * img + 1] = 0; }
    MW_LAP(1);
    if (n_idx < 2) { MW_END(0); return; } // :59-60, nothing is written

    const int pad = (int)ceil(max_distance) + 1; // :70 and :20
    const int l0 = idx[0];                        // :66
    int head = 1;                                 // idx[head .. n_idx) is the remaining list
    int Mb[4];                                    // bounding box of the pixels with distmap == 0
#pragma unroll
    for (int k = 0; k < 4; k++) Mb[k] = box[4 * (l0 - 1) + k];
    if (!aliased) {                               // :68 (a no-op when labels_out is labels)
        const int bw = Mb[3] - Mb[2] + 1, bn = (Mb[1] - Mb[0] + 1) * bw;
        for (int q = tid; q < bn; q += MW_CTA) {
            const int yy = q / bw, p = (Mb[0] + yy) * W + Mb[2] + (q - yy * bw);
            if (L[p] == l0) O[p] = l0;
        }
    }
    if (tid == 0) index_state[2 * img + 1] = 1;

    int win[4];
    win[0] = max(0, Mb[0] - pad); win[1] = (int)min((i64)H, (i64)Mb[1] + 1 + pad);
    win[2] = max(0, Mb[2] - pad); win[3] = (int)min((i64)W, (i64)Mb[3] + 1 + pad);
    // Can anything merge at all?  Search the neighbourhood (radius ceil(max_distance)) of every pixel of another label
    // inside the window for a pixel of l0 -- these pixels are few, the distance map of l0 is the expensive part of a
    // vignette.  If no pair comes within max_distance, every label's minimum of distmap exceeds max_distance^2 (inside
    // the window the distances are exact, outside the map holds its maximum >= pad^2), so whichever label the first
    // iteration pops, merge_dist > max_distance (see the loop) and :94-96 ends the loop before anything is written.
    {
        const double t2d = max_distance * max_distance * (1.0 + 1e-9);
        const int T = t2d >= 2.0e9 ? 0x7fffffff : (int)floor(t2d);
        const int r = pad - 1;
        // (at most ~4 M pixel tests: the budget shrinks with the search radius)
        const i64 side = 2 * (i64)r + 1;
        const int budget = (int)max((i64)1, min((i64)MW_NEAR_BUDGET, ((i64)1 << 22) / (side * side)));
        if (tid == 0) { S.near = 0; S.tested = 0; }
        __syncthreads();
        const int ww = win[3] - win[2], npw = (win[1] - win[0]) * ww;
        for (int q0 = tid; q0 < npw; q0 += 8 * MW_CTA) {
            int lv[8], ys[8], xs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int q = q0 + u * MW_CTA;
                const int yy = q / ww;
                ys[u] = win[0] + yy; xs[u] = win[2] + (q - yy * ww);
                lv[u] = q < npw ? L[(i64)ys[u] * W + xs[u]] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (!(lv[u] > 0 && lv[u] <= bound && lv[u] != l0)) continue;
                if (*(volatile int *)&S.near) continue;
                const int ya = max(Mb[0], ys[u] - r), yb = min(Mb[1], ys[u] + r);
                const int xa = max(Mb[2], xs[u] - r), xb = min(Mb[3], xs[u] + r);
                if (ya > yb || xa > xb) continue; // farther than r from the box of l0
                {
                    // the closest pair of two pixel sets joins BOUNDARY pixels: a pixel whose four neighbours carry the
                    // same label cannot be nearer to l0 than all of them (a large second object costs its outline only)
                    const int y = ys[u], x = xs[u];
                    const int32_t *c = L + (i64)y * W + x;
                    if (y > 0 && y < H - 1 && x > 0 && x < W - 1 && c[-W] == lv[u] && c[W] == lv[u] && c[-1] == lv[u] &&
                        c[1] == lv[u])
                        continue;
                }
                if (atomicAdd(&S.tested, 1) >= budget) { S.near = 1; continue; } // too many: build the map
                bool hit = false;
                for (int y = ya; y <= yb && !hit; y++) {
                    const int dy = y - ys[u];
                    for (int x = xa; x <= xb; x++) {
                        const int dx = x - xs[u];
                        if (dy * dy + dx * dx <= T && L[(i64)y * W + x] == l0) { hit = true; break; }
                    }
                }
                if (hit) S.near = 1;
            }
        }
        __syncthreads();
        if (!S.near) {
            if (tid == 0) index_state[2 * img + 1] = 2; // l0 and the label the first iteration pops
            MW_END(2);
            return;
        }
    }
    MW_LAP(4);
    uint32_t maxd2; // :74 distmap.max()
    {
        const int npw = (win[1] - win[0]) * (win[3] - win[2]);
        const bool sm = npw <= MW_WCAP;
        maxd2 = (uint32_t)mw_edt(S, L, W, l0, win, sm ? sG : G + (i64)win[0] * W + win[2], sm ? win[3] - win[2] : W,
                                 A + (i64)win[0] * W + win[2], W);
    }
    MW_LAP(2);
    // :24 result = full(dist_sliced.max()); result[slices] = dist_sliced: outside the window A still holds MW_BIG
    // (k_mw_prepare), read as min(A, cap) with cap = max_dist from the start
    // per-label minimum of distmap (:83): outside the window every pixel holds max_dist, the `initial`
    {
        const int ww = win[3] - win[2], npw = (win[1] - win[0]) * ww;
        for (int q0 = tid; q0 < npw; q0 += 8 * MW_CTA) {
            int lv[8];
            i64 pp[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int q = q0 + u * MW_CTA;
                const int yy = q / ww;
                pp[u] = (i64)(win[0] + yy) * W + win[2] + (q - yy * ww);
                lv[u] = q < npw ? L[pp[u]] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (lv[u] > 0 && lv[u] <= bound && lv[u] != l0) atomicMin(mintab + (lv[u] - 1), (uint32_t)A[pp[u]]);
        }
    }
    uint32_t cap = maxd2; // distmap = min(A, cap)
    __syncthreads();
    MW_LAP(3);

    int nm = 0;
    while (head < n_idx) { // :81
        // labels that lost pixels to the last bridge: box and minimum again, over the old box
        if (aliased && S.n_dirty > 0) {
            const int nd = S.n_dirty;
            __syncthreads();
            if (nd > MW_DIRTY) {
                for (int j = tid; j < bound; j += MW_CTA) {
                    box[4 * j] = 0x7fffffff; box[4 * j + 1] = -1; box[4 * j + 2] = 0x7fffffff; box[4 * j + 3] = -1;
                    mintab[j] = 0xffffffffu;
                    dflag[j] = 0;
                }
                __syncthreads();
                mw_scan_all<true>(L, A, H, W, bound, box, mintab);
            } else {
                for (int k = 0; k < nd; k++) {
                    const int dl = S.dirty[k];
                    int ob[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) ob[j] = box[4 * (dl - 1) + j];
                    __syncthreads();
                    if (tid == 0) {
                        int *b = box + 4 * (dl - 1);
                        b[0] = 0x7fffffff; b[1] = -1; b[2] = 0x7fffffff; b[3] = -1;
                        mintab[dl - 1] = 0xffffffffu;
                        dflag[dl - 1] = 0;
                    }
                    __syncthreads();
                    for (int y = ob[0] + warp; y <= ob[1]; y += MW_NW) {
                        for (int x0 = ob[2]; x0 <= ob[3]; x0 += 32) {
                            const int x = x0 + lane;
                            const bool ok = x <= ob[3] && L[(i64)y * W + x] == dl;
                            const unsigned grp = __ballot_sync(FULL, ok);
                            if (!grp) continue;
                            uint32_t a = ok ? (uint32_t)A[(i64)y * W + x] : 0xffffffffu;
                            a = __reduce_min_sync(FULL, a);
                            if (lane == 0) {
                                int *b = box + 4 * (dl - 1);
                                atomicMin(b + 0, y);
                                atomicMax(b + 1, y);
                                atomicMin(b + 2, x0 + __ffs(grp) - 1);
                                atomicMax(b + 3, x0 + 31 - __clz(grp));
                                atomicMin(mintab + (dl - 1), a);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) S.n_dirty = 0;
            __syncthreads();
        }
        MW_LAP(4);
        // :83-84 first minimum of min(distmap over the label, max_dist)
        u64 best = ~0ull;
        for (int j = head + tid; j < n_idx; j += MW_CTA) {
            const int l = idx[j];
            // (a label that lost all its pixels has no distmap values: `initial` alone)
            const uint32_t m = box[4 * (l - 1) + 1] < 0 ? maxd2 : min(min(mintab[l - 1], cap), maxd2);
            const u64 key = ((u64)m << 32) | (uint32_t)j;
            best = key < best ? key : best;
        }
        best = mw_min_u64(S, best);
        const int pos = (int)(best & 0xffffffffu);
        const int cur_l = idx[pos];
        __syncthreads();
        for (int j0 = pos; j0 > head; j0 -= MW_CTA) { // pop: the order of the rest is preserved
            const int j = j0 - tid;
            const int t = (j > head) ? idx[j - 1] : 0;
            __syncthreads();
            if (j > head) idx[j] = t;
            __syncthreads();
        }
        if (tid == 0) { idx[head] = cur_l; index_state[2 * img + 1] = head + 1; } // popped entries stay in front, in pop order
        head++;

        MW_LAP(5);
        int cb[4];
#pragma unroll
        for (int k = 0; k < 4; k++) cb[k] = box[4 * (cur_l - 1) + k];
        if (cb[1] < 0) { // :19-20 find_objects returns None for a label without pixels: TypeError
            if (tid == 0) { status[img] = MAZE_ERR_TYPEERROR; n_merge[img] = nm; }
            return;
        }
        // The popped label's minimum of distmap already tells when the loop ends here: if it exceeds max_distance^2
        // (so do the cap and max_dist, which are part of it), then merge_dist > max_distance.  Proof: a pixel p with
        // sqrt(A) + sqrt(B) <= max_distance has A below every fill value, i.e. A(p) = d(p, m)^2 for a merged label m
        // with p in m's window, and lies inside this label's window (outside, B is the window maximum >= pad^2); with
        // x in m and q in this label nearest to p, |x - q| <= max_distance, so q is inside m's window and
        // A(q) <= |x - q|^2 <= max_distance^2 -- but A(q) >= the label's minimum.  (1e-9: |x - q|^2 is an integer, the
        // float64 sum cannot round across max_distance unless max_distance^2 sits that close below an integer.)
        if ((double)(uint32_t)(best >> 32) > max_distance * max_distance * (1.0 + 1e-9)) break; // :94-96 without :87-92
        win[0] = max(0, cb[0] - pad); win[1] = (int)min((i64)H, (i64)cb[1] + 1 + pad);
        win[2] = max(0, cb[2] - pad); win[3] = (int)min((i64)W, (i64)cb[3] + 1 + pad);
        const int wh = win[1] - win[0], ww = win[3] - win[2], npw = wh * ww;
        const bool sm = npw <= MW_WCAP;
        int32_t *Bp = sm ? sB : B + (i64)win[0] * W + win[2];
        const int bs = sm ? ww : W;
        const int fillB = mw_edt(S, L, W, cur_l, win, sm ? sG : G + (i64)win[0] * W + win[2], bs, Bp, bs); // :87
        __syncthreads();
        MW_LAP(6);
        const bool has_out = !(win[0] == 0 && win[1] == H && win[2] == 0 && win[3] == W);
        const bool m_out = Mb[0] < win[0] || Mb[1] >= win[1] || Mb[2] < win[2] || Mb[3] >= win[3];

        // :90-92 min of sqrt(A) + sqrt(B).  sqrt(a) + sqrt(b) >= sqrt(a + b): a pixel whose a + b is clearly above the
        // running minimum squared cannot lower it
        double md = (has_out && m_out) ? sqrt((double)fillB) : INFINITY;
        // Where can the minimum sit?  A pixel of this label that holds the label's minimum m of distmap gives the sum
        // sqrt(m); a pixel with a sum that small has distmap <= m, and if m is below the cap that is an exact distance to
        // a merged label: the pixel lies within sqrt(m) of the box of the merged labels.  (Outside that box distmap >=
        // m + 1, and sqrt(m + 1) > sqrt(m) in float64 too.)
        int sy0 = 0, sy1 = wh, sx0 = 0, sx1 = ww;
        {
            const uint32_t mb = (uint32_t)(best >> 32);
            if (mb < min(cap, maxd2)) {
                const int rr = (int)sqrt((double)mb) + 1;
                sy0 = max(0, Mb[0] - rr - win[0]); sy1 = min(wh, Mb[1] + rr + 1 - win[0]);
                sx0 = max(0, Mb[2] - rr - win[2]); sx1 = min(ww, Mb[3] + rr + 1 - win[2]);
            }
        }
        for (int yy = sy0 + warp; yy < sy1; yy += MW_NW) { // a warp per row, four groups of 32 pixels in flight
            const int32_t *Ar = A + (i64)(win[0] + yy) * W + win[2];
            const int32_t *Br = Bp + (i64)yy * bs;
            for (int xb = sx0; xb < sx1; xb += 128) {
                uint32_t a[4];
                int b[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int xx = xb + 32 * u + lane;
                    const bool in = xx < sx1;
                    a[u] = in ? min((uint32_t)Ar[xx], cap) : 0x3fffffffu;
                    b[u] = in ? Br[xx] : 0x3fffffff;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if ((double)a[u] + (double)b[u] > md * md * (1.0 + 1e-12)) continue;
                    const double s = sqrt((double)a[u]) + sqrt((double)b[u]);
                    md = s < md ? s : md;
                }
            }
        }
        md = mw_min_double(S, md);
        MW_LAP(7);
        if (md > max_distance) break; // :94-96
        const double lim = md + path_tolerance;
        if (tid == 0 && merge_dist) merge_dist[obj0 + nm] = md; // :100
        nm++;
        {
            // :98, :103-106 (labelmap only ever holds l0) and :109-111 inside the window
            const double lim2 = lim * lim;
            for (int yy = warp; yy < wh; yy += MW_NW) { // a warp per row, four groups of 32 pixels in flight
                const int y = win[0] + yy;
                int32_t *Ar = A + (i64)y * W + win[2];
                const int32_t *Br = Bp + (i64)yy * bs;
                const int32_t *Lr = L + (i64)y * W + win[2];
                int32_t *Or = O + (i64)y * W + win[2];
                for (int xb = 0; xb < ww; xb += 128) {
                    uint32_t ar[4];
                    int bv[4], lv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int xx = xb + 32 * u + lane;
                        const bool in = xx < ww;
                        ar[u] = in ? (uint32_t)Ar[xx] : 0u;
                        bv[u] = in ? Br[xx] : 0;
                        lv[u] = in ? Lr[xx] : 0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int xx = xb + 32 * u + lane;
                        if (xx >= ww) continue;
                        const uint32_t araw = ar[u];
                        const uint32_t a = min(araw, cap);
                        const int b = bv[u];
                        const int l = lv[u];
                        const double ab = (double)a + (double)b;
                        bool fill = l == cur_l;
                        if (!fill && !(ab > lim2 * (1.0 + 1e-12))) {
                            if (2.0 * ab < lim2 * (1.0 - 1e-12)) fill = true;
                            else fill = sqrt((double)a) + sqrt((double)b) <= lim;
                        }
                        const uint32_t an = min(a, (uint32_t)b);
                        const bool other = l > 0 && l <= bound && l != l0 && l != cur_l;
                        if (fill) {
                            if (aliased && other) mw_lose_pixel(S, box, mintab, dflag, l, y, win[2] + xx, araw);
                            Or[xx] = l0;
                        }
                        if (an != araw) Ar[xx] = (int32_t)an;
                        if (other && !(fill && aliased)) atomicMin(mintab + (l - 1), an);
                    }
                }
            }
        }
        MW_LAP(8);
        if (has_out && sqrt((double)fillB) <= lim) {
            // the bridge condition can hold outside the window too: exact pass over the rest of the vignette
            const double sfb = sqrt((double)fillB);
            for (int y = warp; y < H; y += MW_NW) {
                const bool inrow = y >= win[0] && y < win[1];
                for (int x = lane; x < W; x += 32) {
                    if (inrow && x >= win[2] && x < win[3]) continue;
                    const i64 p = (i64)y * W + x;
                    const uint32_t araw = (uint32_t)A[p];
                    const uint32_t a = min(araw, cap);
                    if (sqrt((double)a) + sfb <= lim) {
                        const int l = L[p];
                        if (aliased && l > 0 && l <= bound && l != l0 && l != cur_l) mw_lose_pixel(S, box, mintab, dflag, l, y, x, araw);
                        O[p] = l0;
                    }
                }
            }
        }
        if (has_out) cap = min(cap, (uint32_t)fillB); // :109-111 outside the window
        Mb[0] = min(Mb[0], cb[0]); Mb[1] = max(Mb[1], cb[1]); Mb[2] = min(Mb[2], cb[2]); Mb[3] = max(Mb[3], cb[3]);
        __syncthreads();
        MW_LAP(9);
    }
    MW_END(head);
    if (tid == 0) n_merge[img] = nm;
}

int maze_merge_windowed_launch(int n_teams, cudaStream_t s, const int32_t *labels, int32_t *labels_out,
                               const maze_vignette_t *vig, const int32_t *lab_off, int n_obj_cap, double max_distance,
                               double path_tolerance, int32_t *d2a, int32_t *d2b, int32_t *gbuf, int32_t *obj_scratch,
                               double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                               const int32_t *order)
{
    MAZE_CUDA(cudaFuncSetAttribute(k_merge_windowed, cudaFuncAttributeMaxDynamicSharedMemorySize, MW_SMEM_BYTES),
              "k_merge_windowed attr");
    MAZE_CUDA(cudaMemsetAsync(obj_scratch + 2 * (i64)n_obj_cap, 0xff, 4 * (size_t)n_obj_cap * sizeof(int32_t), s),
              "k_mw_prepare memset");
    MAZE_KERNEL(KID_MERGE_PREPARE, s,
                (k_mw_prepare<<<dim3(n_teams, MW_PARTS), MW_PREP_CTA, 0, s>>>(labels, vig, lab_off, n_obj_cap, d2a, obj_scratch,
                                                                             order)));
    MAZE_KERNEL(KID_MERGE_WINDOWED, s,
                (k_merge_windowed<<<n_teams, MW_CTA, MW_SMEM_BYTES, s>>>(labels, labels_out, vig, lab_off, n_obj_cap, max_distance,
                                                                        path_tolerance, d2a, d2b, gbuf, obj_scratch, merge_dist,
                                                                        n_merge, index_state, status, order)));
    return MAZE_OK;
}
