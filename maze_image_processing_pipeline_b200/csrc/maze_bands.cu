// Band pipeline of the LOKI stage: the same chain as maze_fused.cu (threshold -> thresholded-EDT passes ->
// label -> regionprops accumulators), cut so that every CTA is small and uniform and the per-pixel work never
// waits for the per-vignette work.  sm_100a.
//
//   K1 k_band_front   one CTA per BAND (a group of consecutive rows of one vignette, at most
//                     MAZE_BAND_PLANE_WORDS words including the halo rows it recomputes): threshold + pack,
//                     the morphology passes in shared memory, final bit plane -> HBM, the band's horizontal
//                     runs (y, x0, x1) in raster order + per-run intensity statistics -> HBM run list,
//                     per-word run index (for the dense writer).  Touches every pixel once (1 B/px read).
//   K2 k_band_label   one CTA per VIGNETTE, works on the run list only (a few thousand runs per megapixel):
//                     links the runs across rows (and thereby across bands), union-find, raster-order ranking
//                     = the labels of scipy.ndimage.label / skimage.measure.label, per-label accumulators from
//                     the run end points and the per-run statistics, label of every run -> HBM.
//   K3 k_band_write   one CTA per band again: dense outputs (bool mask byte + int32 label per pixel) composed
//                     from bit plane + run labels and streamed out with full 16-byte stores, zeros included
//                     (5 B/px written once, nothing is zero-filled first).  Skipped in COMPACT mode, where the
//                     run list {y, x0, x1, label} (8 B per run) IS the result that crosses PCIe.
//
// Reference behaviour restated (paths relative to the reference root):
//   maze_ipp/loki/pipeline.py:649 / :405      threshold / bool cast
//   maze_ipp/isotropic.py:35-36, 66-67        erosion / dilation compares on the EDT.  scipy's phantom
//                                             background pixel (plane without any 0) is exact for one-band
//                                             vignettes; a multi-band vignette whose pass input has no 0 is
//                                             flagged (fallback) and redone by the per-operator kernels
//   maze_ipp/loki/pipeline.py:430-433         label(): raster-order labels, 8-connectivity
//   maze_ipp/loki/pipeline.py:589-625         per-label RegionProperties reads (accumulators)
#include <stdlib.h>

#include "maze_fused.cuh"

constexpr int PW = MAZE_BAND_PLANE_WORDS;

// run-table classes of the labelling kernel: runs, rows (+2) and bands a vignette may have
#define BAND_T 256
#define BAND_ZB 6144 /* bytes of zeros in shared memory behind the planes (source of the TMA zero fill) */
#ifndef LABEL_T
#define LABEL_T 64
#endif
#define LABEL_MID_T 256 /* (128 threads with 1024 runs: 0.35 ms instead of 0.21 ms for the list; 512 threads: no change) */
#define LABEL_BIG_T 256
#define LABEL_SMALL_CAP 256
#define LABEL_SMALL_HCAP 1026
#define LABEL_SMALL_NB 16
#define LABEL_SMALL_LCAP 8 /* labels with accumulators in shared memory (the others accumulate in HBM) */
#define LABEL_MID_CAP 2048
#define LABEL_MID_HCAP 2050
#define LABEL_MID_NB 128
#define LABEL_LARGE_CAP 4096
#define LABEL_LARGE_HCAP 4098
#define LABEL_LARGE_NB 256
#define LABEL_BIG_CAP 16384
#define LABEL_BIG_HCAP 16386
#define LABEL_BIG_NB 2048

struct BandParams {
    int t_int, n_pass, halo, high_order, stage_cap, do_props, phantom_mask, has_intensity;
    FusedPass pass[4];
};

struct LabelArgs {
    const maze_vignette_t *vig;
    const int32_t *band_off;
    const maze_band_out_t *band_out;
    maze_run_t *runs;
    const maze_run_stat_t *stats;
    int32_t *n_labels, *fallback, *acc_base, *stage_counter;
    u64 *acc_stage;
    double *hi_stage;
    int32_t *ext_stage;
    int32_t *big_list, *big_counter;
    uint8_t *mask; // dense outputs (or NULL): the labelling kernel stores the runs on top of the zero fill
    int32_t *labels;
    int n_img, n_pass, phantom_mask, do_props, high_order, has_intensity, stage_cap;
    long long huge_px; // vignettes with at least this many pixels are labelled by the global-memory kernels
    long long min_area; // label filters (0 / 0: none)
    int clear_border;
    int single_region; // ImageProperties semantics: no labelling, every run belongs to label 1
};

__device__ __forceinline__ void mark_fallback(const LabelArgs &a, int img)
{
    a.fallback[img] = 1;
    a.n_labels[img] = 0;
    a.acc_base[img] = -1;
}

template <int T, int LCAP>
__device__ int label_vignette(const LabelArgs &a, int img, int cap, int hcap, int nbcap, uint32_t *s_mem);

// run table of the labelling that the LAST band CTA of a vignette runs in place (it reuses the CTA's planes)
#define LABEL_INBAND_CAP 4096
#define LABEL_INBAND_HCAP 2050
#define LABEL_INBAND_NB 64

// ---------------------------------------------------------------------------------------------------------
// K1: band front
// ---------------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, 1024 / T) k_band_front(
    const uint8_t *__restrict__ image, const uint8_t *__restrict__ intensity, const maze_vignette_t *__restrict__ vig,
    const maze_band_t *__restrict__ bands, BandParams prm, uint32_t *__restrict__ bits_out,
    maze_run_t *__restrict__ runs, uint32_t *__restrict__ run_pix, maze_run_stat_t *__restrict__ stats,
    int32_t *run_counter, int run_cap, maze_band_out_t *__restrict__ band_out, uint8_t *__restrict__ zmask,
    int32_t *__restrict__ zlabels, LabelArgs la, int32_t *band_done)
{
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ int s_warp[34];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    const maze_band_t bd = bands[blockIdx.x];
    const maze_vignette_t v = vig[bd.img];
    const int H = v.h, W = v.w, wpr = v.wpr;
    const bool whole = bd.y0 == 0 && bd.y1 == H;  // the band is the vignette: flags are exact, no halo
    const int halo = whole ? 0 : prm.halo;
    const int ya = max(0, bd.y0 - halo), yb = min(H, bd.y1 + halo);
    const int Hb = yb - ya, words = Hb * wpr;      // the host guarantees words <= PW
    const int oy0 = bd.y0 - ya, oy1 = bd.y1 - ya;  // own rows in plane coordinates
    uint32_t *A = s_mem, *B = s_mem + words;
    const int step_y = T / wpr, step_k = T - step_y * wpr;
    const int y_first = tid / wpr, k_first = tid - y_first * wpr;
    int zflags = 0;

    // ---- 0. dense outputs: zero fill of the band's share of the mask and label image (the 16-pixel groups of the
    // flat vignette whose first pixel lies in the band's rows) by the TMA engine: ONE thread issues bulk copies
    // from a zeroed shared-memory buffer (cp.async.bulk shared -> global), in slices between the compute phases.
    // No store instruction, no LSU queue: the compute warps never see the 5 bytes per pixel that leave the SM.
    uint8_t *zb = (uint8_t *)(s_mem + 2 * PW);
    const bool zfill = zlabels != nullptr;
    constexpr int ZF_PARTS = 4;
    int zf_next = 0;
    const int zg_lo = (bd.y0 * W + 15) >> 4, zg_n = ((bd.y1 * W + 15) >> 4) - zg_lo;
    if (zfill) {
        for (int i = tid; i < BAND_ZB / 16; i += T) ((uint4 *)zb)[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    auto tma_zero = [&](int upto) { // call after a barrier that follows the zeroing of zb
        if (!zfill || tid != 0) return;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(zb);
        for (; zf_next < upto; zf_next++) {
            const i64 g0 = zg_lo + (i64)zg_n * zf_next / ZF_PARTS, g1 = zg_lo + (i64)zg_n * (zf_next + 1) / ZF_PARTS;
            uint8_t *pl = (uint8_t *)(zlabels + v.pix_off) + 64 * g0, *pm = zmask + v.pix_off + 16 * g0;
            for (i64 left = 64 * (g1 - g0); left > 0; left -= BAND_ZB, pl += BAND_ZB) {
                const uint32_t sz = (uint32_t)(left < BAND_ZB ? left : BAND_ZB);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pl), "r"(src), "r"(sz) : "memory");
            }
            for (i64 left = 16 * (g1 - g0); left > 0; left -= BAND_ZB, pm += BAND_ZB) {
                const uint32_t sz = (uint32_t)(left < BAND_ZB ? left : BAND_ZB);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pm), "r"(src), "r"(sz) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    };

    // ---- 1. threshold + pack of rows [ya, yb) (loki/pipeline.py:649) ------------------------------------
    uint32_t *T0 = (prm.n_pass & 1) ? B : A; // an odd number of passes must start in B to end in A
    bool hz = false;
    {
        const uint8_t *base = image + v.pix_off + (size_t)ya * W;
        const int t = prm.t_int;
        const uint32_t addc = (uint32_t)((t < 128 ? 127 - t : 255 - t) & 0x7f) * 0x01010101u;
        const bool lowmode = t < 128;
        const uint32_t force = t < 0 ? FULL : 0u, kill = t >= 255 ? 0u : FULL;
        const uint32_t inv0 = (prm.n_pass > 0 && prm.pass[0].invert) ? FULL : 0u;
        int y = y_first, k = k_first;
        for (int w = tid; w < words; w += 2 * T) {
            uint32_t raw[2][9], al[2], vmk[2];
            int yu[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const bool on = w + u * T < words;
                const int nvalid = min(32, W - 32 * k);
                const uint8_t *p = base + (size_t)y * W + 32 * k;
                al[u] = (uint32_t)((uintptr_t)p & 3u);
                const uint32_t *q = (const uint32_t *)(p - al[u]);
                const int last = on ? (int)(al[u] + nvalid - 1) >> 2 : -1; // aligned word holding the last pixel
                vmk[u] = valid_mask(W, k);
                yu[u] = y;
#pragma unroll
                for (int i = 0; i < 9; i++) raw[u][i] = (i <= last) ? __ldg(q + i) : 0u;
                y += step_y; k += step_k;
                if (k >= wpr) { k -= wpr; y++; }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (w + u * T < words) {
                    const uint32_t word = lowmode ? threshold32<true>(raw[u], al[u], addc) : threshold32<false>(raw[u], al[u], addc);
                    const uint32_t o = ((word | force) & kill) & vmk[u];
                    T0[w + u * T] = o;
                    if (yu[u] >= oy0 && yu[u] < oy1) hz |= ((o ^ inv0) | ~vmk[u]) != FULL;
                }
            }
        }
    }
    const int hz0 = __syncthreads_or(hz);
    tma_zero(1);
    zflags |= hz0 ? 1 : 0;
    bool phantom = whole && !hz0 && prm.n_pass > 0 && prm.pass[0].use_phantom;

    // ---- 2. thresholded-EDT passes on the band plane (isotropic.py:35-36, 66-67) ------------------------
    // rows outside [ya, yb) read as ones: right at the image border, and harmless at a band border because the
    // halo (sum of the pass radii) absorbs the error before it reaches the band's own rows
    uint32_t *src = T0, *dst = (T0 == A) ? B : A;
    for (int ps = 0; ps < prm.n_pass; ps++) {
        const int R = prm.pass[ps].R;
        const uint32_t inv = prm.pass[ps].invert ? FULL : 0u;
        const uint32_t inv_next = (ps + 1 < prm.n_pass && prm.pass[ps + 1].invert) ? FULL : 0u;
        hz = false;
        if (R >= 0 && R <= 3) {
            const int nstrip = max(1, min(Hb, T / wpr)); // at most T work items: one round
            const int S = (Hb + nstrip - 1) / nstrip;
            const int *wt = prm.pass[ps].w;
            const int pat = wt[0] != R ? -1 : R * 1000 + wt[0] * 100 + (R >= 1 ? wt[1] * 10 : 0) + (R >= 2 ? wt[2] : 0);
            switch (pat) {
            case 1100: morph_columns<1, T>(src, dst, Hb, W, wpr, WFixed<1, 0, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz, oy0, oy1); break;
            case 1110: morph_columns<1, T>(src, dst, Hb, W, wpr, WFixed<1, 1, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz, oy0, oy1); break;
            case 2210: morph_columns<2, T>(src, dst, Hb, W, wpr, WFixed<2, 1, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz, oy0, oy1); break;
            case 2221: morph_columns<2, T>(src, dst, Hb, W, wpr, WFixed<2, 2, 1, 0>(), inv, phantom, nstrip, S, inv_next, hz, oy0, oy1); break;
            case 2222: morph_columns<2, T>(src, dst, Hb, W, wpr, WFixed<2, 2, 2, 0>(), inv, phantom, nstrip, S, inv_next, hz, oy0, oy1); break;
            default:
                if (R == 0) morph_columns<0, T>(src, dst, Hb, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz, oy0, oy1);
                else if (R == 1) morph_columns<1, T>(src, dst, Hb, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz, oy0, oy1);
                else if (R == 2) morph_columns<2, T>(src, dst, Hb, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz, oy0, oy1);
                else morph_columns<3, T>(src, dst, Hb, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz, oy0, oy1);
            }
        } else {
            int y = y_first, k = k_first;
            for (int w = tid; w < words; w += T) {
                uint32_t acc = FULL;
                for (int dy = -R; dy <= R; dy++) {
                    const int yy = y + dy;
                    const int hw = prm.pass[ps].w[dy < 0 ? -dy : dy];
                    const uint32_t C = smem_plane_load(src, Hb, W, wpr, yy, k, inv, phantom);
                    uint32_t h = C;
                    if (hw > 0) {
                        const uint32_t L = smem_plane_load(src, Hb, W, wpr, yy, k - 1, inv, phantom);
                        const uint32_t Rw = smem_plane_load(src, Hb, W, wpr, yy, k + 1, inv, phantom);
                        for (int d = 1; d <= hw; d++) {
                            h &= __funnelshift_rc(C, Rw, d);
                            h &= __funnelshift_lc(L, C, d);
                        }
                    }
                    acc &= h;
                }
                const uint32_t vm = valid_mask(W, k), o = (acc ^ inv) & vm;
                dst[w] = o;
                if (y >= oy0 && y < oy1) hz |= ((o ^ inv_next) | ~vm) != FULL;
                y += step_y; k += step_k;
                if (k >= wpr) { k -= wpr; y++; }
            }
        }
        const int hzp = __syncthreads_or(hz);
        tma_zero(min(2 + ps, ZF_PARTS - 1));
        zflags |= hzp ? (2 << ps) : 0;
        phantom = whole && !hzp && ps + 1 < prm.n_pass && prm.pass[ps + 1].use_phantom;
        uint32_t *tmp = src; src = dst; dst = tmp;
    }
    tma_zero(ZF_PARTS);
    // final plane (== A): own rows -> HBM
    const uint32_t *Mo = src + oy0 * wpr;
    const int n_own = (oy1 - oy0) * wpr;
    {
        uint32_t *gb = bits_out + v.word_off + (size_t)bd.y0 * wpr;
        for (int w = tid; w < n_own; w += T) gb[w] = Mo[w];
    }

    // ---- 3. run list of the own rows -------------------------------------------------------------------
    const int RC = min((4 * (2 * PW - words)) / 6, 65535); // everything behind the final plane is free
    u16 *rY = (u16 *)(s_mem + words), *rX0 = rY + RC, *rX1 = rX0 + RC;
    // every thread walks a contiguous chunk of words; run STARTS and run ENDS are emitted independently (the j-th
    // start and the j-th end of the band belong to the same run), so no thread ever follows a run across words
    const int chunk = ((n_own + T - 1) / T) | 1;
    const int lo = min(tid * chunk, n_own), hi = min(lo + chunk, n_own);
    int cnt = 0, open = 0;
    {
        int k = lo % wpr;
        uint32_t prev = (lo < hi && k > 0) ? Mo[lo - 1] : 0u;
        if (lo < hi && k > 0) open = (int)((prev >> 31) & Mo[lo] & 1u); // a run of the previous chunk continues here
        for (int w = lo; w < hi; w++) {
            const uint32_t m = Mo[w];
            cnt += __popc(m & ~((m << 1) | (k > 0 ? prev >> 31 : 0u)));
            prev = m;
            if (++k == wpr) k = 0;
        }
    }
    int n_runs;
    int run = block_exclusive_scan<T>(cnt, s_warp, &n_runs);
    const bool fits = n_runs <= RC; // else: more runs than slots (noise), the vignette falls back to the per-operator kernels
    if (fits && lo < hi) {
        int y = lo / wpr, k = lo - y * wpr;
        int rs = run, re = run - open; // ends before this chunk = starts before it - runs still open
        uint32_t prev = k > 0 ? Mo[lo - 1] : 0u, m = Mo[lo];
        for (int w = lo; w < hi; w++) {
            const uint32_t nxt = (w + 1 < n_own) ? Mo[w + 1] : 0u;
            const uint32_t pb = k > 0 ? prev >> 31 : 0u, nbit = (k + 1 < wpr) ? (nxt & 1u) : 0u;
            uint32_t starts = m & ~((m << 1) | pb), ends = m & ~((m >> 1) | (nbit << 31));
            const int xw = 32 * k, yy = bd.y0 + y;
            while (starts) {
                const int b0 = __ffs(starts) - 1;
                starts &= starts - 1;
                rY[rs] = (u16)yy; rX0[rs] = (u16)(xw + b0);
                rs++;
            }
            while (ends) {
                const int b1 = __ffs(ends) - 1;
                ends &= ends - 1;
                rX1[re++] = (u16)(xw + b1);
            }
            prev = m; m = nxt;
            if (++k == wpr) { k = 0; y++; }
        }
    }
    if (tid == 0) {
        int base = -1;
        if (fits) {
            base = n_runs ? atomicAdd(run_counter, n_runs) : 0;
            if (base + n_runs > run_cap) base = -1; // run buffer full (the counter still tells the host how many were needed)
        }
        s_base = base;
        band_out[blockIdx.x] = maze_band_out_t{base, n_runs, zflags, 0};
    }
    __syncthreads();
    const int base = s_base;
    if (base >= 0)
    for (int i = tid; i < n_runs; i += T) {
        uint2 r;
        r.x = (uint32_t)rY[i] | ((uint32_t)rX0[i] << 16);
        r.y = (uint32_t)rX1[i];
        ((uint2 *)runs)[base + i] = r;
        if (run_pix) run_pix[base + i] = (uint32_t)(v.pix_off + (i64)rY[i] * W + rX0[i]); // where the dense writer starts
    }
    // ---- 4. intensity statistics per run: four lanes per run (aligned words round robin, two loads in flight per
    // lane, byte-SIMD), reduced across the quad; the bytes were read by this CTA a moment ago, so the loads hit L2 --
    if (intensity && base >= 0) {
        const uint8_t *gi = intensity + v.pix_off;
        const int q = tid & 3;
        for (int i0 = 0; i0 < n_runs; i0 += T / 4) {
            const int i = i0 + (tid >> 2);
            uint32_t sV = 0, sZ = 0, mn = FULL, mx = 0u;
            if (i < n_runs) {
                const int y = rY[i], a = rX0[i], b = rX1[i];
                const uint8_t *prow = gi + (size_t)y * W;
                const int al = (int)((uintptr_t)prow & 3u);
                const uint32_t *qq = (const uint32_t *)(prow - al);
                const int ga = (a + al) >> 2, gb = (b + al) >> 2;
                for (int g0 = ga + q; g0 <= gb; g0 += 8) {
                    uint32_t pxs[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) pxs[u] = __ldg(qq + min(g0 + 4 * u, gb));
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int g = g0 + 4 * u;
                        const uint32_t px = pxs[u];
                        const int c0 = 4 * g - al; // column of byte 0 of this word
                        uint32_t nib = g <= gb ? 0xfu : 0u;
                        if (c0 < a) nib &= 0xfu << (a - c0);
                        if (c0 + 3 > b) nib &= 0xfu >> min(c0 + 3 - b, 31);
                        const uint32_t bm = ((nib * 0x00204081u) & 0x01010101u) * 0xffu;
                        sV = __dp4a(px & bm, 0x01010101u, sV);
                        sZ += __popc(~(((px & 0x7f7f7f7fu) + 0x7f7f7f7fu) | px) & bm & 0x80808080u);
                        mn = __vminu4(mn, px | ~bm);
                        mx = __vmaxu4(mx, px & bm);
                    }
                }
            }
#pragma unroll
            for (int d = 1; d < 4; d <<= 1) {
                sV += __shfl_xor_sync(FULL, sV, d);
                sZ += __shfl_xor_sync(FULL, sZ, d);
                mn = __vminu4(mn, __shfl_xor_sync(FULL, mn, d));
                mx = __vmaxu4(mx, __shfl_xor_sync(FULL, mx, d));
            }
            if (q == 0 && i < n_runs) {
                mn = __vminu4(mn, mn >> 16); mn = __vminu4(mn, mn >> 8);
                mx = __vmaxu4(mx, mx >> 16); mx = __vmaxu4(mx, mx >> 8);
                uint2 st;
                st.x = sV;
                st.y = (sZ & 0xffffu) | ((mn & 0xffu) << 16) | ((mx & 0xffu) << 24);
                ((uint2 *)stats)[base + i] = st;
            }
        }
    }
    if (zfill && tid == 0) { // the fill is complete; order it (async proxy) before the generic stores that follow it
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    // ---- 5. the LAST band of a vignette to get here labels it in place (run-list labelling, accumulators, dense runs;
    // see label_vignette): the latency-bound work on the run list runs inside this memory-bound kernel instead of
    // behind it.  Runs, statistics and the zero fill of every band are complete and visible before its ticket.
    if (band_done) {
        __shared__ int s_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const int nb_v = la.band_off[bd.img + 1] - la.band_off[bd.img];
            s_last = atomicAdd(band_done + bd.img, 1) == nb_v - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (label_vignette<T, FUSED_LCAP>(la, bd.img, LABEL_INBAND_CAP, LABEL_INBAND_HCAP, LABEL_INBAND_NB, s_mem) == 1 && tid == 0)
                la.big_list[2 * la.n_img + atomicAdd(la.big_counter + 2, 1)] = bd.img; // larger than the in-place table: the list kernel takes it
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: per-vignette labelling + accumulators on the run list
// ---------------------------------------------------------------------------------------------------------
// returns 0 = done (labelled, flagged as fallback, or not a band vignette), 1 = needs a larger run table
template <int T, int LCAP>
__device__ int label_vignette(const LabelArgs &a, int img, int cap, int hcap, int nbcap, uint32_t *s_mem)
{
    __shared__ int s_warp[34];
    __shared__ int s_misc[4];
    __shared__ int s_base;
    __shared__ int s_hist[LCAP + 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = a.band_off[img], b1 = a.band_off[img + 1], nb = b1 - b0;
    if (nb <= 0) return 0; // not a band vignette (the per-operator chain owns its counters)
    const maze_vignette_t v = a.vig[img];
    const int H = v.h;
    if ((i64)v.h * v.w >= a.huge_px) return 0; // a frame: the global-memory labelling kernels take it
    if (nb > nbcap) return 1;
    // band table -> shared memory: s_bbase[j] = first run of band j in the run buffer, s_bpre[j] = runs before band j
    int *s_bbase = (int *)s_mem, *s_bpre = s_bbase + nbcap;
    if (tid < 4) s_misc[tid] = 0;
    __syncthreads();
    {
        int bad = 0, zor = 0;
        for (int j = tid; j < nb; j += T) {
            const int4 oq = __ldcg((const int4 *)(a.band_out + b0 + j)); // (written by other CTAs of this launch: not through L1)
            const maze_band_out_t o = {oq.x, oq.y, oq.z, oq.w};
            s_bbase[j] = o.base; s_bpre[j + 1] = o.n_runs;
            bad |= (o.base < 0) ? 1 : 0; zor |= o.zflags;
        }
        bad = __reduce_or_sync(FULL, bad); zor = __reduce_or_sync(FULL, zor);
        if (lane == 0 && (bad | zor)) { atomicOr(&s_misc[1], bad); atomicOr(&s_misc[2], zor); }
    }
    __syncthreads();
    if (warp == 0) { // prefix sum of the run counts (a handful of bands)
        int carry = 0;
        for (int j0 = 0; j0 < nb; j0 += 32) {
            const int j = j0 + lane;
            int x = j < nb ? s_bpre[j + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, x, d);
                if (lane >= d) x += t;
            }
            if (j < nb) s_bpre[j + 1] = carry + x;
            carry += __shfl_sync(FULL, x, 31);
        }
        if (lane == 0) { s_bpre[0] = 0; s_misc[0] = carry; }
    }
    __syncthreads();
    const int n_runs = s_misc[0];
    int bad = s_misc[1];
    const int zor = s_misc[2];
    // position of run i (vignette raster order) in the run buffer
    auto gidx = [&](int i) {
        int lo = 0, hi = nb - 1; // last band j with s_bpre[j] <= i
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_bpre[mid] <= i) lo = mid; else hi = mid - 1;
        }
        return s_bbase[lo] + (i - s_bpre[lo]);
    };
    // scipy's phantom pixel applies to a pass whose (inverted) input plane has no 0 anywhere: a one-band vignette
    // handled it exactly in K1, a multi-band one is redone by the per-operator kernels
    if (nb > 1)
        for (int p = 0; p < a.n_pass; p++)
            if (((a.phantom_mask >> p) & 1) && !((zor >> p) & 1)) bad = 1;
    if (bad) {
        if (tid == 0) mark_fallback(a, img);
        return 0;
    }
    if (a.single_region && n_runs > cap && cap < LABEL_BIG_CAP) return 1; // (the tiers only spread the work here)
    if (a.single_region) {
        // ---- ImageProperties(mask, image), loki/pipeline.py:653: the whole mask of the vignette is ONE region with
        // label 1 (also when it is empty: the row then has area 0 and the caller drops the vignette, :651).  No
        // labelling, hence no run table and no limit on the number of runs: the runs are streamed from the run buffer,
        // sums per thread, one reduction per warp into the single accumulator row.
        AccRow *ACC1 = (AccRow *)(((uintptr_t)(s_bpre + nbcap + 1) + 15) & ~(uintptr_t)15);
        if (tid == 0) {
            a.n_labels[img] = 1;
            a.fallback[img] = 0;
            int base = -1;
            if (a.do_props) {
                base = atomicAdd(a.stage_counter, 1);
                if (base + 1 > a.stage_cap) base = -1;
            }
            s_base = base;
            a.acc_base[img] = base;
        }
        for (int i = tid; i < (int)(sizeof(AccRow) / 4); i += T) ((uint32_t *)ACC1)[i] = 0;
        __syncthreads();
        const int base = s_base;
        const bool props = a.do_props && base >= 0;
        const int W = v.w;
        if (tid == 0) {
            ACC1->e[E_RMIN] = 0x7fffffff; ACC1->e[E_RMAX] = -1; ACC1->e[E_CMIN] = 0x7fffffff; ACC1->e[E_CMAX] = -1;
            ACC1->e[E_VMIN] = 0x7fffffff; ACC1->e[E_VMAX] = -1;
        }
        __syncthreads();
        u64 aN = 0, aR = 0, aC = 0, aRR = 0, aRC = 0, aCC = 0, aRRR = 0, aRRC = 0, aRCC = 0, aCCC = 0, aV = 0, aZ = 0;
        int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1, vmin = 255, vmax = 0;
#pragma unroll 4
        for (int i = tid; i < (props ? n_runs : 0); i += T) { // (independent iterations: several loads in flight)
            const int g = gidx(i);
            const uint2 r = __ldcg((const uint2 *)a.runs + g);
            const u64 y = r.x & 0xffffu, xa = r.x >> 16, xb = r.y & 0xffffu, n = xb - xa + 1;
            const u64 S1 = n * (xa + xb) / 2;
            const u64 S2 = pw2(xb) - (xa ? pw2(xa - 1) : 0);
            const u64 S3 = pw3(xb) - (xa ? pw3(xa - 1) : 0);
            aN += n; aR += y * n; aC += S1; aRR += y * y * n; aRC += y * S1; aCC += S2;
            aRRR += y * y * y * n; aRRC += y * y * S1; aRCC += y * S2; aCCC += S3;
            rmin = min(rmin, (int)y); rmax = max(rmax, (int)y); cmin = min(cmin, (int)xa); cmax = max(cmax, (int)xb);
            if (a.has_intensity) {
                const uint2 st = __ldcg((const uint2 *)a.stats + g);
                aV += st.x; aZ += st.y & 0xffffu;
                vmin = min(vmin, (int)((st.y >> 16) & 0xffu)); vmax = max(vmax, (int)(st.y >> 24));
            }
        }
        if (props) {
            u64 vv[12] = {aN, aR, aC, aRR, aRC, aCC, aRRR, aRRC, aRCC, aCCC, aV, aZ};
#pragma unroll
            for (int j = 0; j < 12; j++) {
                u64 x = vv[j];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                vv[j] = x;
            }
            rmin = __reduce_min_sync(FULL, rmin); rmax = __reduce_max_sync(FULL, rmax);
            cmin = __reduce_min_sync(FULL, cmin); cmax = __reduce_max_sync(FULL, cmax);
            vmin = __reduce_min_sync(FULL, vmin); vmax = __reduce_max_sync(FULL, vmax);
            if (lane == 0 && vv[0]) { // (the label field of the run records stays 0: there is no label image here)
#pragma unroll
                for (int j = 0; j < 10; j++) shared_add64(ACC1->a + j, vv[j]);
                if (a.has_intensity) { shared_add64(ACC1->a + A_V, vv[10]); shared_add64(ACC1->a + A_Z, vv[11]); }
                atomicMin(&ACC1->e[E_RMIN], rmin); atomicMax(&ACC1->e[E_RMAX], rmax);
                atomicMin(&ACC1->e[E_CMIN], cmin); atomicMax(&ACC1->e[E_CMAX], cmax);
                if (a.has_intensity) { atomicMin(&ACC1->e[E_VMIN], vmin); atomicMax(&ACC1->e[E_VMAX], vmax); }
            }
        }
        // dense outputs: mask bytes (and label 1) of every run, eight lanes per run as below
        if (a.labels) {
            int32_t *gl = a.labels + v.pix_off;
            uint8_t *gm = a.mask + v.pix_off;
            const int oct = lane >> 3, ol = lane & 7;
            for (int i0 = warp * 4; i0 < n_runs; i0 += (T / 32) * 4) {
                const int i = i0 + oct;
                if (i < n_runs) {
                    const uint2 r = __ldcg((const uint2 *)a.runs + gidx(i));
                    const int len = (int)(r.y & 0xffffu) - (int)(r.x >> 16) + 1;
                    const int p = (int)(r.x & 0xffffu) * W + (int)(r.x >> 16);
                    int32_t *pl = gl + p;
                    uint8_t *pm = gm + p;
                    const int head = min((-p) & 3, len), body = (len - head) & ~3, tail = len - head - body;
                    if (ol < head) { pl[ol] = 1; pm[ol] = 1; }
                    if (ol >= 4 && ol - 4 < tail) { pl[head + body + ol - 4] = 1; pm[head + body + ol - 4] = 1; }
                    for (int x = head + 4 * ol; x < head + body; x += 32) {
                        *(uint4 *)(pl + x) = make_uint4(1, 1, 1, 1);
                        *(uint32_t *)(pm + x) = 0x01010101u;
                    }
                }
            }
        }
        if (!props) return 0;
        __syncthreads();
        if (a.high_order) { // float64 moments with p + q > 3 about the exact centroid: a second pass over the runs
            const double dnA = (double)*(const volatile u64 *)(ACC1->a + A_N);
            const double cr = (double)*(const volatile u64 *)(ACC1->a + A_R) / dnA, cc = (double)*(const volatile u64 *)(ACC1->a + A_C) / dnA;
            double h13 = 0, h22 = 0, h31 = 0, h23 = 0, h32 = 0, h33 = 0;
#pragma unroll 4
            for (int i = tid; i < n_runs; i += T) {
                const uint2 r = __ldcg((const uint2 *)a.runs + gidx(i));
                const int x0 = r.x >> 16, nn = (int)(r.y & 0xffffu) - x0 + 1;
                const double dn = (double)nn, m1 = (double)(nn - 1);
                const double S1 = dn * m1 * 0.5, S2 = m1 * dn * (2.0 * m1 + 1.0) / 6.0, S3 = S1 * S1;
                const double o = cc - (double)x0;
                const double T1 = S1 - dn * o, T2 = S2 - 2.0 * o * S1 + dn * o * o;
                const double T3 = S3 - 3.0 * o * S2 + 3.0 * o * o * S1 - dn * o * o * o;
                const double dr = (double)(r.x & 0xffffu) - cr, dr2 = dr * dr, dr3 = dr2 * dr;
                h13 += dr * T3; h22 += dr2 * T2; h31 += dr3 * T1;
                h23 += dr2 * T3; h32 += dr3 * T2; h33 += dr3 * T3;
            }
            double hv[6] = {h13, h22, h31, h23, h32, h33};
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double x = hv[j];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                hv[j] = x;
            }
            if (lane == 0 && n_runs > 0) {
                atomicAdd(ACC1->h + H_13, hv[0]); atomicAdd(ACC1->h + H_22, hv[1]); atomicAdd(ACC1->h + H_31, hv[2]);
                atomicAdd(ACC1->h + H_23, hv[3]); atomicAdd(ACC1->h + H_32, hv[4]); atomicAdd(ACC1->h + H_33, hv[5]);
            }
            __syncthreads();
        }
        for (int j = tid; j < MAZE_NACC; j += T) a.acc_stage[(i64)base * MAZE_NACC + j] = ACC1->a[j];
        for (int j = tid; j < 8; j += T) {
            a.hi_stage[(i64)base * 8 + j] = ACC1->h[j];
            a.ext_stage[(i64)base * MAZE_NEXT + j] = ACC1->e[j];
        }
        return 0;
    }
    if (n_runs > cap || n_runs >= 0x8000 || H + 2 > hcap) return 1;

    u16 *rY = (u16 *)(s_bpre + nbcap + 1), *rX0 = rY + cap, *rX1 = rX0 + cap, *P = rX1 + cap, *rO = P + cap, *rowStart = rO + cap;
    AccRow *ACC = (AccRow *)(((uintptr_t)(rowStart + hcap) + 15) & ~(uintptr_t)15);
    for (int i = tid; i < n_runs; i += T) {
        const uint2 r = __ldcg((const uint2 *)a.runs + gidx(i));
        rY[i] = (u16)(r.x & 0xffffu); rX0[i] = (u16)(r.x >> 16); rX1[i] = (u16)(r.y & 0xffffu);
        P[i] = (u16)i;
    }
    __syncthreads();
    int n_lab;
    {
        // first run of every row (runs are in raster order)
        for (int i = tid; i < n_runs; i += T) {
            const int y = rY[i], yp = i ? (int)rY[i - 1] : -1;
            for (int yy = yp + 1; yy <= y; yy++) rowStart[yy] = (u16)i;
        }
        {
            const int ylast = n_runs ? (int)rY[n_runs - 1] : -1;
            for (int yy = ylast + 1 + tid; yy <= H; yy += T) rowStart[yy] = (u16)n_runs;
        }
        __syncthreads();
        // link every run to the runs of the previous row it touches (8-connectivity: x ranges within 1): plain store to
        // the FIRST run it touches, pointer doubling, lock-free unions for the further runs (see maze_fused.cu)
        for (int i = tid; i < n_runs; i += T) {
            const int y = rY[i];
            int first = 0xffff;
            if (y > 0) {
                int p = rowStart[y - 1];
                const int bnd = rowStart[y];
                const int lo0 = (int)rX0[i] - 1, hi0 = (int)rX1[i] + 1;
                int e = bnd;
                while (p < e) {
                    const int mid = (p + e) >> 1;
                    if ((int)rX1[mid] < lo0) p = mid + 1; else e = mid;
                }
                if (p < bnd && (int)rX0[p] <= hi0) { first = p; P[i] = (u16)p; }
            }
            rO[i] = (u16)first;
        }
        __syncthreads();
        for (int i = tid; i < n_runs; i += T) {
            int p = *(volatile u16 *)(P + i);
            for (;;) {
                const int g = *(volatile u16 *)(P + p);
                if (g == p) break;
                *(volatile u16 *)(P + i) = (u16)g;
                p = g;
            }
        }
        __syncthreads();
        for (int i = tid; i < n_runs; i += T) {
            const int first = rO[i];
            if (first == 0xffff) continue;
            const int bnd = rowStart[rY[i]];
            const int hi0 = (int)rX1[i] + 1;
            for (int j = first + 1; j < bnd && (int)rX0[j] <= hi0; j++) union16(P, i, j);
        }
        __syncthreads();
        for (int r = tid; r < n_runs; r += T) {
            const int root = find16(P, r);
            if (root != r) P[r] = (u16)root;
        }
        __syncthreads();
        // roots in raster order get labels 1..N, stored in their own slot with the top bit set
        const int rchunk = (n_runs + T - 1) / T;
        const int rlo = min(tid * rchunk, n_runs), rhi = min(rlo + rchunk, n_runs);
        int nroot = 0;
        for (int r = rlo; r < rhi; r++) nroot += (P[r] == r);
        int rank = block_exclusive_scan<T>(nroot, s_warp, &n_lab);
        for (int r = rlo; r < rhi; r++)
            if (P[r] == r) P[r] = (u16)(0x8000 | (++rank));
    }
    if (tid == 0) {
        a.n_labels[img] = n_lab;
        a.fallback[img] = 0;
        int base = -1;
        if (a.do_props) {
            base = n_lab ? atomicAdd(a.stage_counter, n_lab) : 0;
            if (base + n_lab > a.stage_cap) base = -1; // staging full: maze_regionprops takes this vignette
        }
        s_base = base;
        a.acc_base[img] = base;
    }
    {
        const int nsm = min(n_lab, LCAP);
        for (int i = tid; i < nsm * (int)(sizeof(AccRow) / 4); i += T) ((uint32_t *)ACC)[i] = 0;
        if (tid <= LCAP + 1) s_hist[tid] = 0;
    }
    __syncthreads(); // labels in P are final
    for (int r = tid; r < n_runs; r += T) { // every slot holds its label now (roots already do and never change)
        const int pr = P[r];
        if (pr < 0x8000) P[r] = P[pr];
    }
    __syncthreads();
    const int base = s_base;
    const bool props = a.do_props && base >= 0;
    if (props) {
        const int nsm = min(n_lab, LCAP);
        for (int l = tid; l < nsm; l += T) {
            ACC[l].e[E_RMIN] = 0x7fffffff; ACC[l].e[E_RMAX] = -1; ACC[l].e[E_CMIN] = 0x7fffffff; ACC[l].e[E_CMAX] = -1;
            ACC[l].e[E_VMIN] = 0x7fffffff; ACC[l].e[E_VMAX] = -1;
        }
        for (int l = LCAP + tid; l < n_lab; l += T) { // labels beyond the shared table accumulate in HBM
            u64 *ga = a.acc_stage + (i64)(base + l) * MAZE_NACC;
            for (int j = 0; j < MAZE_NACC; j++) ga[j] = 0;
            double *gh = a.hi_stage + (i64)(base + l) * 8;
            for (int j = 0; j < 8; j++) gh[j] = 0.0;
            int32_t *ge = a.ext_stage + (i64)(base + l) * MAZE_NEXT;
            ge[E_RMIN] = 0x7fffffff; ge[E_RMAX] = -1; ge[E_CMIN] = 0x7fffffff; ge[E_CMAX] = -1;
            ge[E_VMIN] = 0x7fffffff; ge[E_VMAX] = -1; ge[6] = 0; ge[7] = 0;
        }
        for (int i = tid; i < n_runs; i += T) atomicAdd(&s_hist[min(label_of(P, i) - 1, LCAP)], 1);
    }
    __syncthreads();
    // label of every run -> HBM, per-run statistics -> accumulator rows, dense outputs (after the label filters)
    auto write_out = [&]() {
        // ---- label of every run -> HBM; per-run intensity statistics -> accumulator rows (warp-aggregated) -----------
        {
            const bool with_i = props && a.has_intensity;
            for (int i0 = warp * 32; i0 < n_runs; i0 += T) {
                const int i = i0 + lane;
                const bool valid = i < n_runs;
                int lab = 0;
                uint32_t isum = 0, zeros = 0, vmn = 255, vmx = 0;
                if (valid) {
                    const int g = gidx(i);
                    lab = label_of(P, i);
                    ((u16 *)(a.runs + g))[3] = (u16)lab;
                    if (with_i) {
                        const uint2 st = __ldcg((const uint2 *)a.stats + g);
                        isum = st.x; zeros = st.y & 0xffffu; vmn = (st.y >> 16) & 0xffu; vmx = st.y >> 24;
                    }
                }
                if (with_i) {
                    uint32_t todo = __ballot_sync(FULL, valid);
                    while (todo) {
                        const int leader = __ffs(todo) - 1;
                        const int L = __shfl_sync(FULL, lab, leader);
                        const bool in = valid && lab == L;
                        todo &= ~__ballot_sync(FULL, in);
                        const uint32_t sV = __reduce_add_sync(FULL, in ? isum : 0u), sZ = __reduce_add_sync(FULL, in ? zeros : 0u);
                        const uint32_t mn = __reduce_min_sync(FULL, in ? vmn : 255u), mx = __reduce_max_sync(FULL, in ? vmx : 0u);
                        if (lane == leader) {
                            if (L <= LCAP) {
                                shared_add64(ACC[L - 1].a + A_V, (u64)sV);
                                if (sZ) shared_add64(ACC[L - 1].a + A_Z, (u64)sZ);
                                atomicMin(&ACC[L - 1].e[E_VMIN], (int)mn); atomicMax(&ACC[L - 1].e[E_VMAX], (int)mx);
                            } else {
                                u64 *Aa = a.acc_stage + (i64)(base + L - 1) * MAZE_NACC;
                                int32_t *Ee = a.ext_stage + (i64)(base + L - 1) * MAZE_NEXT;
                                atomicAdd(Aa + A_V, (u64)sV); atomicAdd(Aa + A_Z, (u64)sZ);
                                atomicMin(Ee + E_VMIN, (int)mn); atomicMax(Ee + E_VMAX, (int)mx);
                            }
                        }
                    }
                }
            }
        }
        // ---- dense outputs: the runs of the vignette on top of the zero fill of K1.  Eight lanes per run: the up to three
        // elements in front of the first 16-byte boundary of the label row (lanes 0-2) and behind the last one (lanes
        // 4-6) are stored one by one, the body with one 16-byte label store and one 4-byte mask store per lane and step ---
        if (a.labels) {
            int32_t *gl = a.labels + v.pix_off;
            uint8_t *gm = a.mask + v.pix_off;
            const int W = v.w, oct = lane >> 3, ol = lane & 7;
            for (int i0 = warp * 4; i0 < n_runs; i0 += (T / 32) * 4) {
                const int i = i0 + oct;
                if (i < n_runs) {
                    const int len = (int)rX1[i] - (int)rX0[i] + 1;
                    const int p = (int)rY[i] * W + (int)rX0[i];
                    const uint32_t lab = (uint32_t)label_of(P, i);
                    int32_t *pl = gl + p;
                    uint8_t *pm = gm + p;
                    const int head = min((-p) & 3, len), body = (len - head) & ~3, tail = len - head - body;
                    if (ol < head) { pl[ol] = (int32_t)lab; pm[ol] = 1; }
                    if (ol >= 4 && ol - 4 < tail) { pl[head + body + ol - 4] = (int32_t)lab; pm[head + body + ol - 4] = 1; }
                    const uint4 l4 = make_uint4(lab, lab, lab, lab);
                    for (int x = head + 4 * ol; x < head + body; x += 32) {
                        *(uint4 *)(pl + x) = l4;
                        *(uint32_t *)(pm + x) = 0x01010101u;
                    }
                }
            }
        }
    };
    if (!props) {
        if (a.clear_border || a.min_area > 0) { // the filters need the accumulators: per-operator redo
            if (tid == 0) mark_fallback(a, img);
            return 0;
        }
        write_out();
        return 0;
    }

    // ---- per-label geometry accumulators: counting sort of the runs by label, contiguous pieces per thread ------
    if (tid == 0) {
        int acc0 = 0;
        for (int b = 0; b <= LCAP; b++) { const int c = s_hist[b]; s_hist[b] = acc0; acc0 += c; }
    }
    __syncthreads();
    for (int i = tid; i < n_runs; i += T) rO[atomicAdd(&s_hist[min(label_of(P, i) - 1, LCAP)], 1)] = (u16)i;
    __syncthreads();
    const int p_chunk = (n_runs + T - 1) / T;
    const int p_lo = min(tid * p_chunk, n_runs), p_hi = min(p_lo + p_chunk, n_runs);
    {
        int cur = 0;
        u64 aN = 0, aR = 0, aC = 0, aRR = 0, aRC = 0, aCC = 0, aRRR = 0, aRRC = 0, aRCC = 0, aCCC = 0;
        int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1;
        auto flush = [&](int lab) {
            const u64 sums[10] = {aN, aR, aC, aRR, aRC, aCC, aRRR, aRRC, aRCC, aCCC};
            int *Ee;
            if (lab <= LCAP) {
                u64 *Aa = ACC[lab - 1].a; Ee = ACC[lab - 1].e;
#pragma unroll
                for (int j = 0; j < 10; j++) shared_add64(Aa + j, sums[j]);
            } else {
                u64 *Aa = a.acc_stage + (i64)(base + lab - 1) * MAZE_NACC; Ee = a.ext_stage + (i64)(base + lab - 1) * MAZE_NEXT;
#pragma unroll
                for (int j = 0; j < 10; j++) atomicAdd(Aa + j, sums[j]);
            }
            atomicMin(Ee + E_RMIN, rmin); atomicMax(Ee + E_RMAX, rmax);
            atomicMin(Ee + E_CMIN, cmin); atomicMax(Ee + E_CMAX, cmax);
        };
        for (int pos = p_lo; pos < p_hi; pos++) {
            const int i = rO[pos];
            const int L = label_of(P, i);
            if (L != cur) {
                if (cur) flush(cur);
                aN = aR = aC = aRR = aRC = aCC = aRRR = aRRC = aRCC = aCCC = 0;
                rmin = 0x7fffffff; rmax = -1; cmin = 0x7fffffff; cmax = -1;
                cur = L;
            }
            const u64 y = rY[i], xa = rX0[i], xb = rX1[i], n = xb - xa + 1;
            const u64 S1 = n * (xa + xb) / 2;
            const u64 S2 = pw2(xb) - (xa ? pw2(xa - 1) : 0);
            const u64 S3 = pw3(xb) - (xa ? pw3(xa - 1) : 0);
            aN += n; aR += y * n; aC += S1; aRR += y * y * n; aRC += y * S1; aCC += S2;
            aRRR += y * y * y * n; aRRC += y * y * S1; aRCC += y * S2; aCCC += S3;
            rmin = min(rmin, (int)y); rmax = max(rmax, (int)y); cmin = min(cmin, (int)xa); cmax = max(cmax, (int)xb);
        }
        uint32_t todo = __ballot_sync(FULL, cur != 0);
        while (todo) { // one flush per label and warp
            const int leader = __ffs(todo) - 1;
            const int lab = __shfl_sync(FULL, cur, leader);
            const bool in = (cur == lab);
            todo &= ~__ballot_sync(FULL, in);
            u64 vv[10] = {aN, aR, aC, aRR, aRC, aCC, aRRR, aRRC, aRCC, aCCC};
#pragma unroll
            for (int j = 0; j < 10; j++) {
                u64 x = in ? vv[j] : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                vv[j] = x;
            }
            const int r0 = __reduce_min_sync(FULL, in ? rmin : 0x7fffffff), r1 = __reduce_max_sync(FULL, in ? rmax : -1);
            const int c0 = __reduce_min_sync(FULL, in ? cmin : 0x7fffffff), c1 = __reduce_max_sync(FULL, in ? cmax : -1);
            if (lane == leader) {
                aN = vv[0]; aR = vv[1]; aC = vv[2]; aRR = vv[3]; aRC = vv[4]; aCC = vv[5]; aRRR = vv[6]; aRRC = vv[7];
                aRCC = vv[8]; aCCC = vv[9]; rmin = r0; rmax = r1; cmin = c0; cmax = c1;
                flush(lab);
            }
        }
    }
    __syncthreads();
    // ---- label filters on the run list (loki/pipeline.py:435-448): clear_border removes every label that touches the
    // outermost rows / columns, remove_small_objects every label with fewer than min_area pixels.  Both follow from
    // the accumulators just taken (bbox, area): the runs of a removed label get label 0 (no renumbering, as in the
    // reference; their mask bytes stay 1) and its row is emptied.
    if (a.clear_border || a.min_area > 0) {
        const int W = v.w;
        __threadfence();
        auto removed = [&](int lab) {
            const u64 *Aa = lab <= LCAP ? ACC[lab - 1].a : a.acc_stage + (i64)(base + lab - 1) * MAZE_NACC;
            const int *Ee = lab <= LCAP ? ACC[lab - 1].e : a.ext_stage + (i64)(base + lab - 1) * MAZE_NEXT;
            if (a.min_area > 0 && (i64) * (const volatile u64 *)(Aa + A_N) < a.min_area) return true;
            if (a.clear_border)
                return *(const volatile int *)(Ee + E_RMIN) == 0 || *(const volatile int *)(Ee + E_RMAX) == H - 1 ||
                       *(const volatile int *)(Ee + E_CMIN) == 0 || *(const volatile int *)(Ee + E_CMAX) == W - 1;
            return false;
        };
        for (int i = tid; i < n_runs; i += T)
            if (removed(label_of(P, i))) P[i] = (u16)0x8000; // label 0
        __syncthreads();
        for (int l = tid; l < n_lab; l += T)
            if (removed(l + 1)) {
                u64 *Aa = l < LCAP ? ACC[l].a : a.acc_stage + (i64)(base + l) * MAZE_NACC;
                int *Ee = l < LCAP ? ACC[l].e : a.ext_stage + (i64)(base + l) * MAZE_NEXT;
                for (int j = 0; j < MAZE_NACC; j++) Aa[j] = 0;
                Ee[E_RMIN] = 0x7fffffff; Ee[E_RMAX] = -1; Ee[E_CMIN] = 0x7fffffff; Ee[E_CMAX] = -1;
            }
        __syncthreads();
    }
    write_out();
    __syncthreads();
    if (a.high_order) {
        // float64 central moments with p + q > 3 about the exact centroid (same walk over the sorted runs)
        for (int l = tid; l < n_lab; l += T) {
            const u64 *Aa = l < LCAP ? ACC[l].a : a.acc_stage + (i64)(base + l) * MAZE_NACC;
            double *Hh = l < LCAP ? ACC[l].h : a.hi_stage + (i64)(base + l) * 8;
            const double dn = (double)*(const volatile u64 *)(Aa + A_N);
            Hh[H_CR] = (double)*(const volatile u64 *)(Aa + A_R) / dn;
            Hh[H_CC] = (double)*(const volatile u64 *)(Aa + A_C) / dn;
        }
        __syncthreads();
        int cur = 0;
        double h13 = 0, h22 = 0, h31 = 0, h23 = 0, h32 = 0, h33 = 0, cr = 0, cc = 0;
        auto hrow = [&](int lab) { return lab <= LCAP ? ACC[lab - 1].h : a.hi_stage + (i64)(base + lab - 1) * 8; };
        for (int pos = p_lo; pos < p_hi; pos++) {
            const int i = rO[pos];
            const int L = label_of(P, i);
            if (L == 0) continue; // removed by a label filter
            if (L != cur) {
                if (cur) {
                    double *Hh = hrow(cur);
                    atomicAdd(Hh + H_13, h13); atomicAdd(Hh + H_22, h22); atomicAdd(Hh + H_31, h31);
                    atomicAdd(Hh + H_23, h23); atomicAdd(Hh + H_32, h32); atomicAdd(Hh + H_33, h33);
                }
                h13 = h22 = h31 = h23 = h32 = h33 = 0;
                cur = L;
                const double *Hc = hrow(L);
                cr = *(const volatile double *)(Hc + H_CR);
                cc = *(const volatile double *)(Hc + H_CC);
            }
            const int nn = (int)rX1[i] - (int)rX0[i] + 1;
            const double dn = (double)nn, m1 = (double)(nn - 1);
            const double S1 = dn * m1 * 0.5;
            const double S2 = m1 * dn * (2.0 * m1 + 1.0) / 6.0;
            const double S3 = S1 * S1;
            const double o = cc - (double)rX0[i];
            const double T1 = S1 - dn * o;
            const double T2 = S2 - 2.0 * o * S1 + dn * o * o;
            const double T3 = S3 - 3.0 * o * S2 + 3.0 * o * o * S1 - dn * o * o * o;
            const double dr = (double)rY[i] - cr, dr2 = dr * dr, dr3 = dr2 * dr;
            h13 += dr * T3; h22 += dr2 * T2; h31 += dr3 * T1;
            h23 += dr2 * T3; h32 += dr3 * T2; h33 += dr3 * T3;
        }
        uint32_t todo = __ballot_sync(FULL, cur != 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int lab = __shfl_sync(FULL, cur, leader);
            const bool in = (cur == lab);
            todo &= ~__ballot_sync(FULL, in);
            double vv[6] = {h13, h22, h31, h23, h32, h33};
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double x = in ? vv[j] : 0.0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                vv[j] = x;
            }
            if (lane == leader) {
                double *Hh = hrow(lab);
                atomicAdd(Hh + H_13, vv[0]); atomicAdd(Hh + H_22, vv[1]); atomicAdd(Hh + H_31, vv[2]);
                atomicAdd(Hh + H_23, vv[3]); atomicAdd(Hh + H_32, vv[4]); atomicAdd(Hh + H_33, vv[5]);
            }
        }
        __syncthreads();
    }
    {   // shared rows -> staging
        const int nsm = min(n_lab, LCAP);
        for (int i = tid; i < nsm * MAZE_NACC; i += T) {
            const int l = i / MAZE_NACC, j = i - l * MAZE_NACC;
            a.acc_stage[(i64)(base + l) * MAZE_NACC + j] = ACC[l].a[j];
        }
        for (int i = tid; i < nsm * 8; i += T) {
            const int l = i >> 3, j = i & 7;
            a.hi_stage[(i64)(base + l) * 8 + j] = ACC[l].h[j];
            a.ext_stage[(i64)(base + l) * MAZE_NEXT + j] = ACC[l].e[j];
        }
    }
    return 0;
}

// small class: one CTA per vignette; a vignette with more runs (rows, bands) than the class holds goes on the list of
// the smallest class that holds it (list k = big_list[k * n_img ..], its length in big_counter[k], the counter its
// CTAs draw from in big_counter[3 + k])
template <int T>
__global__ void __launch_bounds__(T, 1024 / T) k_band_label(LabelArgs a, int cap, int hcap)
{
    extern __shared__ __align__(16) uint32_t s_mem[];
    const int img = blockIdx.x;
    if (label_vignette<T, LABEL_SMALL_LCAP>(a, img, cap, hcap, LABEL_SMALL_NB, s_mem) == 1 && threadIdx.x == 0) {
        int tot = 0;
        const int nb = a.band_off[img + 1] - a.band_off[img], h2 = a.vig[img].h + 2;
        for (int b = a.band_off[img]; b < a.band_off[img + 1]; b++) tot += a.band_out[b].n_runs;
        const int k = (tot <= LABEL_MID_CAP && h2 <= LABEL_MID_HCAP && nb <= LABEL_MID_NB) ? 0
                      : (tot <= LABEL_LARGE_CAP && h2 <= LABEL_LARGE_HCAP && nb <= LABEL_LARGE_NB) ? 1 : 2;
        a.big_list[k * a.n_img + atomicAdd(a.big_counter + k, 1)] = img;
    }
}

// the list classes: a fixed grid, the CTAs take the vignettes of their list one by one from a counter
template <int T>
__global__ void __launch_bounds__(T, 1024 / T) k_band_label_big(LabelArgs a, int cap, int hcap, int nbcap, int which)
{
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ int s_next;
    const int n = *(volatile int32_t *)(a.big_counter + which);
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(a.big_counter + 3 + which, 1);
        __syncthreads();
        const int e = s_next;
        __syncthreads();
        if (e >= n) break;
        const int img = a.big_list[which * a.n_img + e];
        if (label_vignette<T, FUSED_LCAP>(a, img, cap, hcap, nbcap, s_mem) == 1 && threadIdx.x == 0) mark_fallback(a, img);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: dense writer -- every run of the batch's run buffer is stored on top of the zero fill (a warp fetches 32
// run records at once, then stores run after run with coalesced 4-byte label / 1-byte mask stores).  Runs of
// vignettes that were flagged as fallback carry label 0 and are overwritten by the per-operator kernels later.
// ---------------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T) k_band_write(const maze_run_t *__restrict__ runs, const uint32_t *__restrict__ run_pix,
                                                  const int32_t *__restrict__ run_counter, int run_cap,
                                                  uint8_t *__restrict__ mask, int32_t *__restrict__ labels)
{
    const int n = min(*run_counter, run_cap);
    const int lane = threadIdx.x & 31, oct = lane >> 3, ol = lane & 7; // four octets per warp, one run each at a time
    const int nwarp = gridDim.x * (T / 32);
    for (int i0 = (blockIdx.x * (T / 32) + (threadIdx.x >> 5)) * 32; i0 < n; i0 += nwarp * 32) {
        const int nr = min(32, n - i0);
        uint32_t pix = 0, rec_x = 0, rec_y = 0;
        if (lane < nr) {
            const uint2 r = ((const uint2 *)runs)[i0 + lane];
            pix = run_pix[i0 + lane]; rec_x = r.x; rec_y = r.y;
        }
        const uint32_t len_me = lane < nr ? (rec_y & 0xffffu) - (rec_x >> 16) + 1u : 0u;
        // an octet stores 32 elements per step, aligned to 32 elements: a whole 32-byte sector of the mask and a
        // whole 128-byte line of the label image per step (but for the two ends of the run); a lane owns 4 elements
        for (int j = oct; j < 32; j += 4) { // (all lanes take part in the shuffles; runs past nr have length 0)
            const uint32_t p = __shfl_sync(FULL, pix, j);
            const int len = (int)__shfl_sync(FULL, len_me, j);
            const uint32_t lab = __shfl_sync(FULL, rec_y, j) >> 16;
            int32_t *gl = labels + p;
            uint8_t *gm = mask + p;
            for (int x = 4 * ol - (int)(p & 31u); x < len; x += 32) {
                if (x >= 0 && x + 4 <= len) {
                    *(uint4 *)(gl + x) = make_uint4(lab, lab, lab, lab);
                    *(uint32_t *)(gm + x) = 0x01010101u;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (x + u >= 0 && x + u < len) { gl[x + u] = (int32_t)lab; gm[x + u] = 1; }
                }
            }
        }
    }
}

// zero fill of the dense outputs (runs on a forked stream next to the band front, which leaves the memory system idle)
__global__ void __launch_bounds__(256) k_band_zero(uint4 *__restrict__ a, size_t na, uint4 *__restrict__ b, size_t nb)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) a[i] = z;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride) b[i] = z;
}

// the same fill by the TMA engine: every lane streams 16 KB bulk copies from a zeroed shared-memory buffer
// (cp.async.bulk shared -> global, SASS UBLKCP): no store instructions, no LSU queue, 32 threads per SM
#define ZCH 16384
__global__ void __launch_bounds__(32) k_band_zero_tma(uint8_t *__restrict__ a, size_t na, uint8_t *__restrict__ b, size_t nb)
{
    extern __shared__ __align__(128) uint8_t zbuf[];
    const int lane = threadIdx.x;
    for (int i = lane; i < ZCH / 16; i += 32) ((uint4 *)zbuf)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
    const size_t ca = (na + ZCH - 1) / ZCH, cb = (nb + ZCH - 1) / ZCH;
    int pending = 0;
    for (size_t c = (size_t)blockIdx.x * 32 + lane; c < ca + cb; c += (size_t)gridDim.x * 32) {
        uint8_t *dst;
        size_t left;
        if (c < ca) { dst = a + c * ZCH; left = na - c * ZCH; }
        else { dst = b + (c - ca) * ZCH; left = nb - (c - ca) * ZCH; }
        const uint32_t size = (uint32_t)(left < (size_t)ZCH ? left : (size_t)ZCH); // multiples of 16 bytes
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(size) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++pending == 8) {
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            pending = 4;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int zero_mode()
{
    static int m = -1;
    if (m < 0) {
        const char *e = getenv("MAZE_ZERO_MODE");
        m = !e ? 2 : e[0] == 't' ? 1 : e[0] == 's' ? 0 : 2; // default: inside the band front (TMA)
    }
    return m;
}

static int zero_ctas()
{
    static int n = 0;
    if (!n) {
        const char *e = getenv("MAZE_ZERO_CTAS"); // experiments: throttle the fill so that it leaves DRAM headroom
        n = e ? atoi(e) : 148 * 4;
        if (n < 1) n = 1;
    }
    return n;
}

struct BandFork {
    int device;
    cudaStream_t aux;
    cudaEvent_t fork, join;
};

static BandFork *band_fork()
{
    static thread_local BandFork pool[16];
    static thread_local int n_pool = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < n_pool; i++)
        if (pool[i].device == dev) return &pool[i];
    if (n_pool >= 16) return nullptr;
    BandFork *f = &pool[n_pool];
    f->device = dev;
    if (cudaStreamCreateWithFlags(&f->aux, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    n_pool++;
    return f;
}

// experiment (MAZE_K2_PRIO=1): the labelling kernels on a HIGH-PRIORITY stream, so that their CTAs take the slots the
// band front of the next lane frees, instead of queueing behind it
static BandFork *band_fork_prio()
{
    static thread_local BandFork pool[16];
    static thread_local int n_pool = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < n_pool; i++)
        if (pool[i].device == dev) return &pool[i];
    if (n_pool >= 16) return nullptr;
    BandFork *f = &pool[n_pool];
    f->device = dev;
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&f->aux, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    n_pool++;
    return f;
}

// ---------------------------------------------------------------------------------------------------------
// K2 for FRAMES (BASELINE.json configs[3]: 4096 x 4096, thousands of labels, 10^5 runs): the same labelling and
// accumulation on the run list, but with the union-find in global memory and every step spread over the whole GPU.
// One vignette per call sequence: prefix of the band run counts -> init -> link (atomicMin union-find on run ids)
// -> flatten -> rank (raster order of the roots = label numbers) -> zero rows -> apply (labels into the run records,
// integer accumulators by atomics) -> high-order moments -> dense write.
// scratch (int32): parent[run_cap] | lab[run_cap] | bpre[n_bands_of_vignette + 1] | misc[8]
// ---------------------------------------------------------------------------------------------------------
struct GlArgs {
    LabelArgs a;
    int img;
    int32_t *parent, *lab, *bpre, *misc; // misc: [0] runs, [1] ok, [2] labels, [3] staging base
};

__device__ __forceinline__ int gl_band_of(const int32_t *bpre, int nb, int i)
{
    int lo = 0, hi = nb - 1; // last band j with bpre[j] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (bpre[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(1024) k_gl_prefix(GlArgs g)
{
    __shared__ int s_warp[34];
    const LabelArgs &a = g.a;
    const int b0 = a.band_off[g.img], nb = a.band_off[g.img + 1] - b0;
    const int tid = threadIdx.x;
    const int chunk = (nb + 1023) / 1024, lo = min(tid * chunk, nb), hi = min(lo + chunk, nb);
    int sum = 0, bad = 0, zor = 0;
    for (int j = lo; j < hi; j++) {
        const maze_band_out_t o = a.band_out[b0 + j];
        sum += o.n_runs; bad |= o.base < 0 ? 1 : 0; zor |= o.zflags;
    }
    bad = __syncthreads_or(bad);
    // OR of the zero flags over the bands
    __shared__ int s_zor;
    if (tid == 0) s_zor = 0;
    __syncthreads();
    if (zor) atomicOr(&s_zor, zor);
    int total;
    int run = block_exclusive_scan<1024>(sum, s_warp, &total);
    for (int j = lo; j < hi; j++) {
        g.bpre[j] = run;
        run += a.band_out[b0 + j].n_runs;
    }
    if (tid == 0) {
        g.bpre[nb] = total;
        if (nb > 1)
            for (int p = 0; p < a.n_pass; p++)
                if (((a.phantom_mask >> p) & 1) && !((s_zor >> p) & 1)) bad = 1; // scipy's phantom pixel: per-operator redo
        g.misc[0] = bad ? 0 : total;
        g.misc[1] = bad ? 0 : 1;
        g.misc[2] = 0;
        g.misc[3] = -1;
        if (bad) mark_fallback(a, g.img);
    }
}

__global__ void __launch_bounds__(256) k_gl_init(GlArgs g)
{
    const int n = g.misc[0];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) g.parent[i] = i;
}

__global__ void __launch_bounds__(256) k_gl_link(GlArgs g)
{
    const LabelArgs &a = g.a;
    const int n = g.misc[0];
    const int b0 = a.band_off[g.img], nb = a.band_off[g.img + 1] - b0;
    const uint2 *R = (const uint2 *)a.runs;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int jb = gl_band_of(g.bpre, nb, i);
        const uint2 r = R[a.band_out[b0 + jb].base + (i - g.bpre[jb])];
        const int y = r.x & 0xffffu, x0 = r.x >> 16, x1 = r.y & 0xffffu;
        if (y == 0) continue;
        // runs of row y - 1 live in band pb (the same or the previous one), sorted by x: first one whose end reaches x0 - 1
        int pb = jb;
        if (i == g.bpre[jb] || (int)(R[a.band_out[b0 + jb].base].x & 0xffffu) > y - 1) pb = jb - 1; // band starts at row y
        if (pb < 0) continue;
        const maze_band_out_t po = a.band_out[b0 + pb];
        const uint2 *P = R + po.base;
        int lo = 0, hi = po.n_runs; // first run with (row, x1) >= (y - 1, x0 - 1)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const uint2 q = P[mid];
            const int qy = q.x & 0xffffu, qx1 = q.y & 0xffffu;
            if (qy < y - 1 || (qy == y - 1 && qx1 < x0 - 1)) lo = mid + 1; else hi = mid;
        }
        for (int j = lo; j < po.n_runs; j++) {
            const uint2 q = P[j];
            if ((int)(q.x & 0xffffu) != y - 1 || (int)(q.x >> 16) > x1 + 1) break;
            uf_union(g.parent, i, g.bpre[pb] + j);
        }
    }
}

__global__ void __launch_bounds__(256) k_gl_flatten(GlArgs g)
{
    const int n = g.misc[0];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) g.lab[i] = uf_find(g.parent, i);
}

// roots in raster order get labels 1..N (one CTA: the scan is over run ids, a few 10^5 at most)
__global__ void __launch_bounds__(1024) k_gl_rank(GlArgs g)
{
    __shared__ int s_warp[34];
    const LabelArgs &a = g.a;
    const int n = g.misc[0];
    if (!g.misc[1]) return;
    const int tid = threadIdx.x;
    const int chunk = (n + 1023) / 1024, lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int cnt = 0;
    for (int i = lo; i < hi; i++) cnt += g.lab[i] == i;
    int n_lab;
    int rank = block_exclusive_scan<1024>(cnt, s_warp, &n_lab);
    for (int i = lo; i < hi; i++)
        if (g.lab[i] == i) g.parent[i] = ++rank; // parent[root] = label (roots only; the others keep their root in lab)
    if (tid == 0) {
        if (n_lab > 65535) { // labels are 16 bits in the run records
            mark_fallback(a, g.img);
            g.misc[0] = 0; g.misc[1] = 0;
            return;
        }
        a.n_labels[g.img] = n_lab;
        a.fallback[g.img] = 0;
        int base = -1;
        if (a.do_props) {
            base = n_lab ? atomicAdd(a.stage_counter, n_lab) : 0;
            if (base + n_lab > a.stage_cap) base = -1;
        }
        a.acc_base[g.img] = base;
        g.misc[2] = n_lab;
        g.misc[3] = base;
    }
}

__global__ void __launch_bounds__(256) k_gl_zero_rows(GlArgs g)
{
    const LabelArgs &a = g.a;
    const int n_lab = g.misc[2], base = g.misc[3];
    if (base < 0) return;
    for (int l = blockIdx.x * 256 + threadIdx.x; l < n_lab; l += gridDim.x * 256) {
        u64 *ga = a.acc_stage + (i64)(base + l) * MAZE_NACC;
        for (int j = 0; j < MAZE_NACC; j++) ga[j] = 0;
        double *gh = a.hi_stage + (i64)(base + l) * 8;
        for (int j = 0; j < 8; j++) gh[j] = 0.0;
        int32_t *ge = a.ext_stage + (i64)(base + l) * MAZE_NEXT;
        ge[E_RMIN] = 0x7fffffff; ge[E_RMAX] = -1; ge[E_CMIN] = 0x7fffffff; ge[E_CMAX] = -1;
        ge[E_VMIN] = 0x7fffffff; ge[E_VMAX] = -1; ge[6] = 0; ge[7] = 0;
    }
}

// stage 0: label of every run -> run record, integer accumulators; stage 1: float64 moments about the centroid
__global__ void __launch_bounds__(256) k_gl_apply(GlArgs g, int stage)
{
    const LabelArgs &a = g.a;
    const int n = g.misc[0], base = g.misc[3];
    const int b0 = a.band_off[g.img], nb = a.band_off[g.img + 1] - b0;
    const bool props = a.do_props && base >= 0;
    if (stage == 1 && !(props && a.high_order)) return;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int jb = gl_band_of(g.bpre, nb, i);
        const int gi = a.band_out[b0 + jb].base + (i - g.bpre[jb]);
        const uint2 r = ((const uint2 *)a.runs)[gi];
        const int lab = g.parent[g.lab[i]];
        const u64 y = r.x & 0xffffu, xa = r.x >> 16, xb = r.y & 0xffffu, nn = xb - xa + 1;
        if (stage == 0) {
            ((u16 *)(a.runs + gi))[3] = (u16)lab;
            if (!props) continue;
            u64 *Aa = a.acc_stage + (i64)(base + lab - 1) * MAZE_NACC;
            int32_t *Ee = a.ext_stage + (i64)(base + lab - 1) * MAZE_NEXT;
            const u64 S1 = nn * (xa + xb) / 2;
            const u64 S2 = pw2(xb) - (xa ? pw2(xa - 1) : 0);
            const u64 S3 = pw3(xb) - (xa ? pw3(xa - 1) : 0);
            atomicAdd(Aa + A_N, nn); atomicAdd(Aa + A_R, y * nn); atomicAdd(Aa + A_C, S1);
            atomicAdd(Aa + A_RR, y * y * nn); atomicAdd(Aa + A_RC, y * S1); atomicAdd(Aa + A_CC, S2);
            atomicAdd(Aa + A_RRR, y * y * y * nn); atomicAdd(Aa + A_RRC, y * y * S1); atomicAdd(Aa + A_RCC, y * S2);
            atomicAdd(Aa + A_CCC, S3);
            atomicMin(Ee + E_RMIN, (int)y); atomicMax(Ee + E_RMAX, (int)y);
            atomicMin(Ee + E_CMIN, (int)xa); atomicMax(Ee + E_CMAX, (int)xb);
            if (a.has_intensity) {
                const uint2 st = ((const uint2 *)a.stats)[gi];
                atomicAdd(Aa + A_V, (u64)st.x);
                if (st.y & 0xffffu) atomicAdd(Aa + A_Z, (u64)(st.y & 0xffffu));
                atomicMin(Ee + E_VMIN, (int)((st.y >> 16) & 0xffu)); atomicMax(Ee + E_VMAX, (int)(st.y >> 24));
            }
        } else {
            const u64 *Aa = a.acc_stage + (i64)(base + lab - 1) * MAZE_NACC;
            double *Hh = a.hi_stage + (i64)(base + lab - 1) * 8;
            const double area = (double)Aa[A_N], cr = (double)Aa[A_R] / area, cc = (double)Aa[A_C] / area;
            const double dn = (double)nn, m1 = (double)(nn - 1);
            const double S1 = dn * m1 * 0.5, S2 = m1 * dn * (2.0 * m1 + 1.0) / 6.0, S3 = S1 * S1;
            const double o = cc - (double)xa;
            const double T1 = S1 - dn * o, T2 = S2 - 2.0 * o * S1 + dn * o * o;
            const double T3 = S3 - 3.0 * o * S2 + 3.0 * o * o * S1 - dn * o * o * o;
            const double dr = (double)y - cr, dr2 = dr * dr, dr3 = dr2 * dr;
            atomicAdd(Hh + H_13, dr * T3); atomicAdd(Hh + H_22, dr2 * T2); atomicAdd(Hh + H_31, dr3 * T1);
            atomicAdd(Hh + H_23, dr2 * T3); atomicAdd(Hh + H_32, dr3 * T2); atomicAdd(Hh + H_33, dr3 * T3);
        }
    }
}

// dense outputs of the frame: its runs on top of the zero fill (eight lanes per run, as in the labelling kernel)
__global__ void __launch_bounds__(256) k_gl_write(GlArgs g)
{
    const LabelArgs &a = g.a;
    const int n = g.misc[0];
    if (!a.labels) return;
    const int b0 = a.band_off[g.img], nb = a.band_off[g.img + 1] - b0;
    const maze_vignette_t v = a.vig[g.img];
    int32_t *gl = a.labels + v.pix_off;
    uint8_t *gm = a.mask + v.pix_off;
    const int W = v.w, lane = threadIdx.x & 31, oct = lane >> 3, ol = lane & 7;
    const int nwarp = gridDim.x * 8;
    for (int i0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4; i0 < n; i0 += nwarp * 4) {
        const int i = i0 + oct;
        if (i >= n) continue;
        const int jb = gl_band_of(g.bpre, nb, i);
        const uint2 r = ((const uint2 *)a.runs)[a.band_out[b0 + jb].base + (i - g.bpre[jb])];
        const int len = (int)(r.y & 0xffffu) - (int)(r.x >> 16) + 1;
        const i64 p = (i64)(r.x & 0xffffu) * W + (r.x >> 16);
        const uint32_t lab = r.y >> 16;
        int32_t *pl = gl + p;
        uint8_t *pm = gm + p;
        for (int x = 4 * ol - (int)(p & 31); x < len; x += 32) {
            if (x >= 0 && x + 4 <= len) {
                *(uint4 *)(pl + x) = make_uint4(lab, lab, lab, lab);
                *(uint32_t *)(pm + x) = 0x01010101u;
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (x + u >= 0 && x + u < len) { pl[x + u] = (int32_t)lab; pm[x + u] = 1; }
            }
        }
    }
}

static int gl_label_frame(const LabelArgs &la, int img, int32_t *scratch, int run_cap, int nb, cudaStream_t s)
{
    GlArgs g;
    g.a = la;
    g.img = img;
    g.parent = scratch;
    g.lab = scratch + run_cap;
    g.bpre = scratch + 2 * (size_t)run_cap;
    g.misc = g.bpre + nb + 1;
    const int grid = 148 * 4;
    MAZE_KERNEL(KID_GL_PREFIX, s, k_gl_prefix<<<1, 1024, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_LINK, s, k_gl_init<<<grid, 256, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_LINK, s, k_gl_link<<<grid, 256, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_LINK, s, k_gl_flatten<<<grid, 256, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_RANK, s, k_gl_rank<<<1, 1024, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_APPLY, s, k_gl_zero_rows<<<grid, 256, 0, s>>>(g));
    MAZE_KERNEL(KID_GL_APPLY, s, k_gl_apply<<<grid, 256, 0, s>>>(g, 0));
    MAZE_KERNEL(KID_GL_APPLY, s, k_gl_apply<<<grid, 256, 0, s>>>(g, 1));
    MAZE_KERNEL(KID_GL_APPLY, s, k_gl_write<<<grid, 256, 0, s>>>(g));
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static size_t label_smem(int cap, int hcap, int nbcap, int lcap)
{
    return (size_t)(2 * nbcap + 1) * 4 + (size_t)cap * 10 + (size_t)hcap * 2 + 16 + lcap * sizeof(AccRow);
}

extern "C" int maze_band_stage(const uint8_t *image, const uint8_t *intensity, const maze_vignette_t *vig, int n_img,
                               const maze_band_t *bands, int n_bands, const int32_t *band_off, int t_int, int n_pass,
                               const int32_t *pass_t_host, const int32_t *pass_invert_host, int halo, int flags,
                               uint32_t *bits, maze_run_t *runs, uint32_t *run_pix, maze_run_stat_t *run_stats, int run_cap,
                               maze_band_out_t *band_out, uint8_t *mask, int32_t *labels, int32_t *n_labels,
                               int32_t *fallback, int32_t *acc_base, int32_t *counters, int32_t *big_list,
                               int stage_cap, unsigned long long *acc_stage, double *hi_stage, int32_t *ext_stage,
                               long long total_px, const int32_t *huge_host, int n_huge, long long huge_px,
                               int32_t *gl_scratch, int32_t *band_done, int clear_border, long long min_area,
                               void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pass < 0 || n_pass > 4 || halo < 0) return MAZE_ERR_BADARG;
    if (n_img <= 0 || n_bands <= 0) return MAZE_OK;
    BandParams prm;
    prm.t_int = t_int;
    prm.n_pass = n_pass;
    prm.halo = halo;
    prm.high_order = (flags & MAZE_RP_HIGH_ORDER) ? 1 : 0;
    prm.stage_cap = stage_cap;
    prm.do_props = (flags & MAZE_FUSED_NO_PROPS) ? 0 : 1;
    prm.has_intensity = intensity ? 1 : 0;
    prm.phantom_mask = 0;
    int sum_r = 0;
    for (int p = 0; p < 4; p++) {
        prm.pass[p].R = -1;
        prm.pass[p].invert = 0;
        prm.pass[p].use_phantom = 1;
        for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) prm.pass[p].w[i] = 0;
    }
    for (int p = 0; p < n_pass; p++) {
        prm.pass[p].invert = pass_invert_host[p] ? 1 : 0;
        if (!maze_pass_table(pass_t_host[p], &prm.pass[p].R, prm.pass[p].w, &prm.pass[p].use_phantom)) return MAZE_ERR_BADARG;
        if (prm.pass[p].use_phantom) prm.phantom_mask |= 1 << p;
        sum_r += prm.pass[p].R > 0 ? prm.pass[p].R : 0;
    }
    if (halo < sum_r) return MAZE_ERR_BADARG; // the halo must absorb every pass
    const bool dense = mask && labels;
    MAZE_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(int32_t), s), "band counters");
    int32_t *stage_counter = counters, *run_counter = counters + 1, *big_counter = counters + 2;
    static thread_local int attr_dev = -1;
    int dev = 0;
    MAZE_CUDA(cudaGetDevice(&dev), "get device");
    static const size_t smem_pad = getenv("MAZE_K1_PAD") ? (size_t)atoi(getenv("MAZE_K1_PAD")) : 0; // experiments
    const size_t smem1 = (size_t)PW * 8 + BAND_ZB + smem_pad, smem_s = label_smem(LABEL_SMALL_CAP, LABEL_SMALL_HCAP, LABEL_SMALL_NB, LABEL_SMALL_LCAP),
                 smem_m = label_smem(LABEL_MID_CAP, LABEL_MID_HCAP, LABEL_MID_NB, FUSED_LCAP),
                 smem_l = label_smem(LABEL_LARGE_CAP, LABEL_LARGE_HCAP, LABEL_LARGE_NB, FUSED_LCAP),
                 smem_b = label_smem(LABEL_BIG_CAP, LABEL_BIG_HCAP, LABEL_BIG_NB, FUSED_LCAP);
    if (attr_dev != dev) {
        MAZE_CUDA(cudaFuncSetAttribute(k_band_front<BAND_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1), "band smem");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_front<BAND_T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "band carveout");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label<LABEL_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s), "label smem");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label<LABEL_T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "label carveout");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label_big<LABEL_BIG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b), "label big smem");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label_big<LABEL_MID_T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(LABEL_MID_T == LABEL_BIG_T ? smem_b : smem_m)), "label mid smem");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label_big<LABEL_MID_T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "label mid carveout");
        // kernels only share an SM when they ask for the same shared-memory carve-out: everything that should run
        // next to the band front (zero fill, labelling, writer of another lane) asks for the maximum like it does
        MAZE_CUDA(cudaFuncSetAttribute(k_band_label_big<LABEL_BIG_T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "label big carveout");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_write<BAND_T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "write carveout");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_zero, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "zero carveout");
        MAZE_CUDA(cudaFuncSetAttribute(k_band_zero_tma, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "zero tma carveout");
        attr_dev = dev;
    }
    BandFork *fk = nullptr;
    if (dense && (total_px <= 0 || (total_px & 15) || total_px >= (1ll << 32) || !run_pix)) return MAZE_ERR_BADARG;
    const bool zin = dense && zero_mode() == 2; // zero fill inside the band front
    // experiment (MAZE_K2_FUSED=1): labelling inside the band front, by the last band CTA of every vignette.  Correct
    // (all parity tests), but the band kernel then takes 0.86 ms instead of 0.57 + 0.30 ms for the two kernels: the
    // labelling holds a CTA slot and the issue slots the band front needs, nothing is hidden.  Off by default.
    static const bool inband_env = getenv("MAZE_K2_FUSED") && getenv("MAZE_K2_FUSED")[0] == '1';
    const bool inband = inband_env && band_done && (!dense || zin);
    if (dense && !zin) { // experiments: zero fill by a kernel of its own next to the band front
        fk = band_fork();
        if (!fk) return MAZE_ERR_CUDA;
        MAZE_CUDA(cudaEventRecord(fk->fork, s), "band fork");
        MAZE_CUDA(cudaStreamWaitEvent(fk->aux, fk->fork, 0), "band fork wait");
        if (zero_mode() == 1) {
            const char *e = getenv("MAZE_ZERO_CTAS");
            MAZE_KERNEL(KID_BAND_ZERO, fk->aux,
                        k_band_zero_tma<<<e ? zero_ctas() : 148, 32, ZCH, fk->aux>>>((uint8_t *)labels, (size_t)total_px * 4,
                                                                                   mask, (size_t)total_px));
        } else {
            MAZE_KERNEL(KID_BAND_ZERO, fk->aux,
                        k_band_zero<<<zero_ctas(), 256, 0, fk->aux>>>((uint4 *)labels, (size_t)total_px / 4, (uint4 *)mask,
                                                                  (size_t)total_px / 16));
        }
        MAZE_CUDA(cudaEventRecord(fk->join, fk->aux), "band join");
    }
    static const bool k3 = getenv("MAZE_K3") != nullptr; // experiments: runs stored by a kernel of their own
    LabelArgs la = {vig, band_off, band_out, runs, run_stats, n_labels, fallback, acc_base, stage_counter,
                    (u64 *)acc_stage, hi_stage, ext_stage, big_list, big_counter, dense && !k3 ? mask : nullptr,
                    dense && !k3 ? labels : nullptr, n_img, n_pass, prm.phantom_mask, prm.do_props, prm.high_order,
                    prm.has_intensity, stage_cap, (n_huge > 0 && gl_scratch) ? huge_px : (1ll << 62), min_area,
                    clear_border ? 1 : 0, (flags & MAZE_BAND_SINGLE_REGION) ? 1 : 0};
    if (inband) MAZE_CUDA(cudaMemsetAsync(band_done, 0, sizeof(int32_t) * (size_t)n_img, s), "band done");
    MAZE_KERNEL(KID_BAND_FRONT, s,
                k_band_front<BAND_T><<<n_bands, BAND_T, smem1, s>>>(image, intensity, vig, bands, prm, bits, runs,
                                                                   dense ? run_pix : nullptr, run_stats, run_counter,
                                                                   run_cap, band_out, zin ? mask : nullptr,
                                                                   zin ? labels : nullptr, la, inband ? band_done : nullptr));
    if (dense && fk) MAZE_CUDA(cudaStreamWaitEvent(s, fk->join, 0), "band join wait");
    const int grid_xl = n_img < 74 ? n_img : 74;
    static const bool k2prio = getenv("MAZE_K2_PRIO") && getenv("MAZE_K2_PRIO")[0] == '1';
    BandFork *pk = nullptr;
    const cudaStream_t s_main = s;
    if (k2prio && !inband) {
        pk = band_fork_prio();
        if (!pk) return MAZE_ERR_CUDA;
        MAZE_CUDA(cudaEventRecord(pk->fork, s), "k2 fork");
        MAZE_CUDA(cudaStreamWaitEvent(pk->aux, pk->fork, 0), "k2 fork wait");
        s = pk->aux;
    }
    if (inband) { // what the in-place labelling passed on (more than 4096 runs, 2048 rows or 64 bands)
        MAZE_KERNEL(KID_BAND_LABEL_BIG, s,
                    k_band_label_big<LABEL_BIG_T><<<grid_xl, LABEL_BIG_T, smem_b, s>>>(la, LABEL_BIG_CAP, LABEL_BIG_HCAP,
                                                                                     LABEL_BIG_NB, 2));
    } else {
        MAZE_KERNEL(KID_BAND_LABEL, s,
                    k_band_label<LABEL_T><<<n_img, LABEL_T, smem_s, s>>>(la, LABEL_SMALL_CAP, LABEL_SMALL_HCAP));
        const int grid_mid = n_img < 148 * 7 ? n_img : 148 * 7, grid_large = n_img < 148 * 4 ? n_img : 148 * 4;
        MAZE_KERNEL(KID_BAND_LABEL_BIG, s,
                    k_band_label_big<LABEL_MID_T><<<grid_mid, LABEL_MID_T, smem_m, s>>>(la, LABEL_MID_CAP, LABEL_MID_HCAP,
                                                                                      LABEL_MID_NB, 0));
        MAZE_KERNEL(KID_BAND_LABEL_BIG, s,
                    k_band_label_big<LABEL_BIG_T><<<grid_large, LABEL_BIG_T, smem_l, s>>>(la, LABEL_LARGE_CAP, LABEL_LARGE_HCAP,
                                                                                        LABEL_LARGE_NB, 1));
        MAZE_KERNEL(KID_BAND_LABEL_BIG, s,
                    k_band_label_big<LABEL_BIG_T><<<grid_xl, LABEL_BIG_T, smem_b, s>>>(la, LABEL_BIG_CAP, LABEL_BIG_HCAP,
                                                                                     LABEL_BIG_NB, 2));
    }
    if (pk) {
        MAZE_CUDA(cudaEventRecord(pk->join, s), "k2 join");
        s = s_main;
        MAZE_CUDA(cudaStreamWaitEvent(s, pk->join, 0), "k2 join wait");
    }
    if (n_huge > 0 && gl_scratch) { // frames: labelling in global memory, one after the other
        if (!huge_host || clear_border || min_area > 0 || (flags & MAZE_BAND_SINGLE_REGION)) return MAZE_ERR_BADARG; // (not on this path)
        for (int e = 0; e < n_huge; e++) {
            // (the band count of the vignette sizes its prefix array; the host knows the band plan)
            const int rc = gl_label_frame(la, huge_host[2 * e], gl_scratch, run_cap, huge_host[2 * e + 1], s);
            if (rc != MAZE_OK) return rc;
        }
    }
    if (dense && k3) {
        MAZE_KERNEL(KID_BAND_WRITE, s,
                    k_band_write<BAND_T><<<148 * 8, BAND_T, 0, s>>>(runs, run_pix, run_counter, run_cap, mask, labels));
    }
    return MAZE_OK;
}
