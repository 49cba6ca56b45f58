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

#include "maze_fused.cuh"

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
    __shared__ int s_hist[FUSED_LCAP + 2];
    const int img = img_list[blockIdx.x];
    const maze_vignette_t v = vig[img];
    const int H = v.h, W = v.w, wpr = v.wpr, words = H * wpr;
    // planes are laid out tightly (A = first `words` words, B = the next `words`); the pass sequence starts in
    // the plane that makes the FINAL plane land in A, so that everything behind A is free for the run list
    uint32_t *A = s_mem, *B = s_mem + words;
    AccRow *ACC = (AccRow *)(s_mem + 2 * wcap);
    const int tid = threadIdx.x;

    // ---- 0. zero-fill of the mask and label image, in ZF_PARTS slices issued between the compute phases: the
    // stores drain in the background instead of blocking every warp at the store queue in one burst
    constexpr int ZF_PARTS = 8;
    int zf_next = 0;
    auto zero_fill = [&](int upto) { // issue slices zf_next .. upto-1
        const int npx = H * W;
        const int nl = (npx + 3) / 4, nm = (npx + 15) / 16;
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *l4 = (uint4 *)(labels + v.pix_off);
        uint4 *m4 = (uint4 *)(mask + v.pix_off);
        for (; zf_next < upto; zf_next++) {
            const int l0 = (int)((i64)nl * zf_next / ZF_PARTS), l1 = (int)((i64)nl * (zf_next + 1) / ZF_PARTS);
#pragma unroll 4
            for (int i = l0 + tid; i < l1; i += T) l4[i] = z;
            const int m0 = (int)((i64)nm * zf_next / ZF_PARTS), m1 = (int)((i64)nm * (zf_next + 1) / ZF_PARTS);
#pragma unroll 2
            for (int i = m0 + tid; i < m1; i += T) m4[i] = z;
        }
    };
    zero_fill(1);
    // word walk without divisions: thread t visits words t, t + T, ...; (y, k) advance by (T / wpr, T % wpr)
    const int step_y = T / wpr, step_k = T - step_y * wpr;
    const int y_first = tid / wpr, k_first = tid - y_first * wpr;

    // ---- 1. threshold + pack (loki/pipeline.py:649) -------------------------------------------------
    uint32_t *T0 = (prm.n_pass & 1) ? B : A; // an odd number of passes must start in B to end in A
    bool hz = false;
    {
        const uint8_t *base = image + v.pix_off;
        const int t = prm.t_int;
        // byte-wise "pixel > t" on four pixels with three integer ops instead of the emulated SIMD compare:
        // (threshold32 above)
        const uint32_t addc = (uint32_t)((t < 128 ? 127 - t : 255 - t) & 0x7f) * 0x01010101u;
        const bool lowmode = t < 128;
        const uint32_t force = t < 0 ? FULL : 0u, kill = t >= 255 ? 0u : FULL;
        const uint32_t inv0 = (prm.n_pass > 0 && prm.pass[0].invert) ? FULL : 0u;
        // two words per trip: all 18 loads are issued before the first compare (the phase is bound by HBM latency)
        int y = y_first, k = k_first;
        {
            int yn = y + 2 * step_y, kn = k + 2 * step_k;
            if (kn >= wpr) { kn -= wpr; yn++; }
            if (kn >= wpr) { kn -= wpr; yn++; }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (tid + (2 + u) * T < words)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)yn * W + 32 * kn));
                yn += step_y; kn += step_k;
                if (kn >= wpr) { kn -= wpr; yn++; }
            }
        }
        for (int w = tid; w < words; w += 2 * T) {
            uint32_t raw[2][9], al[2], vmk[2];
            { // L2 prefetch two trips ahead (every 32-byte sector of the image holds one word start)
                int yn = y + 4 * step_y, kn = k + 4 * step_k; // two trips ahead
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (kn >= wpr) { kn -= wpr; yn++; }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (w + (4 + u) * T < words)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)yn * W + 32 * kn));
                    yn += step_y; kn += step_k;
                    if (kn >= wpr) { kn -= wpr; yn++; }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const bool on = w + u * T < words;
                const int nvalid = min(32, W - 32 * k);
                const uint8_t *p = base + (size_t)y * W + 32 * k;
                al[u] = (uint32_t)((uintptr_t)p & 3u);
                const uint32_t *q = (const uint32_t *)(p - al[u]);
                const int last = on ? (int)(al[u] + nvalid - 1) >> 2 : -1; // aligned word holding the last pixel
                vmk[u] = valid_mask(W, k);
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
                    hz |= ((o ^ inv0) | ~vmk[u]) != FULL;  // the first pass's input plane has a 0 (scipy's phantom pixel otherwise)
                }
            }
        }
    }
    zero_fill(2);
    // scipy's phantom background pixel: the (inverted) input plane of a pass has no 0 at all.  Every producer of a
    // plane (threshold, previous pass) tracks that while it writes, so the barrier between phases carries it.
    bool phantom = !__syncthreads_or(hz) && prm.n_pass > 0 && prm.pass[0].use_phantom;

    // ---- 2. thresholded-EDT passes in shared memory (isotropic.py:35-36, 66-67) ----------------------
    uint32_t *src = T0, *dst = (T0 == A) ? B : A;
    for (int ps = 0; ps < prm.n_pass; ps++) {
        const int R = prm.pass[ps].R;
        const uint32_t inv = prm.pass[ps].invert ? FULL : 0u;
        const uint32_t inv_next = (ps + 1 < prm.n_pass && prm.pass[ps + 1].invert) ? FULL : 0u;
        hz = false;
        if (R >= 0 && R <= 3) {
            // strips of S rows x word columns: about one work item per thread
            const int nstrip = max(1, min(H, (T + wpr - 1) / wpr));
            const int S = (H + nstrip - 1) / nstrip;
            const int *wt = prm.pass[ps].w;
            // (registered footprints may have chords wider than R: they take the run-time table)
            const int pat = wt[0] != R ? -1 : R * 1000 + wt[0] * 100 + (R >= 1 ? wt[1] * 10 : 0) + (R >= 2 ? wt[2] : 0);
            switch (pat) { // d2 thresholds 1, 2-3, 4, 5-7, 8 get fully unrolled code
            case 1100: morph_columns<1, T>(src, dst, H, W, wpr, WFixed<1, 0, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz); break;
            case 1110: morph_columns<1, T>(src, dst, H, W, wpr, WFixed<1, 1, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz); break;
            case 2210: morph_columns<2, T>(src, dst, H, W, wpr, WFixed<2, 1, 0, 0>(), inv, phantom, nstrip, S, inv_next, hz); break;
            case 2221: morph_columns<2, T>(src, dst, H, W, wpr, WFixed<2, 2, 1, 0>(), inv, phantom, nstrip, S, inv_next, hz); break;
            case 2222: morph_columns<2, T>(src, dst, H, W, wpr, WFixed<2, 2, 2, 0>(), inv, phantom, nstrip, S, inv_next, hz); break;
            default:
                if (R == 0) morph_columns<0, T>(src, dst, H, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz);
                else if (R == 1) morph_columns<1, T>(src, dst, H, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz);
                else if (R == 2) morph_columns<2, T>(src, dst, H, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz);
                else morph_columns<3, T>(src, dst, H, W, wpr, WRuntime{wt}, inv, phantom, nstrip, S, inv_next, hz);
            }
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
                const uint32_t vm = valid_mask(W, k), o = (acc ^ inv) & vm;
                dst[w] = o;
                hz |= ((o ^ inv_next) | ~vm) != FULL;
                y += step_y; k += step_k;
                if (k >= wpr) { k -= wpr; y++; }
            }
        }
        zero_fill(min(3 + ps, 5));
        phantom = !__syncthreads_or(hz) && ps + 1 < prm.n_pass && prm.pass[ps + 1].use_phantom;
        uint32_t *tmp = src; src = dst; dst = tmp;
    }
    const uint32_t *M = src; // final plane
    // The free plane now holds the RUN LIST: every maximal horizontal run of foreground pixels (y, x0, x1),
    // its union-find slot, the label-sorted order and the first run of every row.  Everything after the
    // morphology works on this compact list (thread per run, no divergence over empty words).
    const int RC = max(0, (4 * (2 * wcap - words) - 2 * (H + 2)) / 10); // M == A: all words behind it are free
    u16 *rY = (u16 *)(s_mem + words), *rX0 = rY + RC, *rX1 = rX0 + RC, *P = rX1 + RC, *rO = P + RC, *rowStart = rO + RC;

    // ---- 3. run list + labelling (loki/pipeline.py:430-433) --------------------------------------------
    const int chunk = ((words + T - 1) / T) | 1;  // odd: lanes start in different banks
    const int lo = min(tid * chunk, words), hi = min(lo + chunk, words);
    int cnt = 0;
    {
        int k = lo % wpr;
        for (int w = lo; w < hi; w++) {
            uint32_t m = M[w];
            uint32_t prevbit = k > 0 ? (M[w - 1] >> 31) : 0u;
            cnt += __popc(m & ~((m << 1) | prevbit));
            if (++k == wpr) k = 0;
        }
    }
    int n_runs;
    int run = block_exclusive_scan<T>(cnt, s_warp, &n_runs);
    if (n_runs > RC || n_runs >= 0x8000) { // more runs than slots (noise): the per-operator kernels take over
        if (tid == 0) { fallback[img] = 1; n_labels[img] = 0; acc_base[img] = -1; }
        return;
    }
    {
        int y = lo / wpr, k = lo - y * wpr;
        for (int w = lo; w < hi; w++) {
            if (k == 0) rowStart[y] = (u16)run;
            uint32_t m = M[w];
            uint32_t prevbit = k > 0 ? (M[w - 1] >> 31) : 0u;
            uint32_t starts = m & ~((m << 1) | prevbit);
            while (starts) {
                int b0 = __ffs(starts) - 1;
                starts &= starts - 1;
                uint32_t rest = ~(m >> b0);
                int x0 = 32 * k + b0, x1;
                if ((rest & (b0 ? ((1u << (32 - b0)) - 1u) : FULL)) != 0u) {
                    x1 = x0 + __ffs(rest) - 2; // the run ends inside this word
                } else {                       // it reaches bit 31: follow it through the next words of the row
                    x1 = 32 * k + 31;
                    for (int kk = k + 1; kk < wpr; kk++) {
                        uint32_t mm = M[w - k + kk];
                        if (mm == FULL) { x1 += 32; continue; }
                        x1 += __ffs(~mm) - 1;
                        break;
                    }
                }
                rY[run] = (u16)y; rX0[run] = (u16)x0; rX1[run] = (u16)x1; P[run] = (u16)run;
                run++;
            }
            if (++k == wpr) { k = 0; y++; }
        }
        if (tid == 0) rowStart[H] = (u16)n_runs;
    }
    zero_fill(6);
    __syncthreads();
    // link every run to the runs of the previous row it touches (8-connectivity: x ranges within 1), in three steps
    // that keep the union-find chains short: (a) a plain store links the run to the FIRST run it touches (nobody
    // else writes that slot yet), (b) every run jumps its own pointer up to its root (concurrent pointer doubling;
    // only the owner writes a slot, and it always writes an ancestor), (c) the further runs it touches -- the places
    // where components merge -- are united with the lock-free union on chains of length one or two.
    for (int i = tid; i < n_runs; i += T) {
        const int y = rY[i];
        int first = 0xffff;
        if (y > 0) {
            int a = rowStart[y - 1];
            const int bnd = rowStart[y];
            const int lo0 = (int)rX0[i] - 1, hi0 = (int)rX1[i] + 1;
            int e = bnd; // first run of the previous row whose end reaches lo0 (runs of a row are sorted by x)
            while (a < e) {
                int mid = (a + e) >> 1;
                if ((int)rX1[mid] < lo0) a = mid + 1; else e = mid;
            }
            if (a < bnd && (int)rX0[a] <= hi0) { first = a; P[i] = (u16)a; }
        }
        rO[i] = (u16)first; // (the sort order is built much later)
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
    zero_fill(7);
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
    zero_fill(ZF_PARTS);
    {
        uint32_t *gb = bits_out + v.word_off;
        for (int w = tid; w < words; w += T) gb[w] = M[w];
        const int nsm = min(n_lab, FUSED_LCAP);
        for (int i = tid; i < nsm * (int)(sizeof(AccRow) / 4); i += T) ((uint32_t *)ACC)[i] = 0;
        if (tid <= FUSED_LCAP + 1) s_hist[tid] = 0;
    }
    __syncthreads(); // labels in P are final; the zero fill is ordered before the stores below
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
        // one warp per run: coalesced 4-byte label / 1-byte mask stores
        for (int i = warp; i < n_runs; i += NWARP) {
            const int y = rY[i], x0 = rX0[i], x1 = rX1[i];
            const int lab = label_of(P, i);
            const int o = y * W;
            for (int x = x0 + lane; x <= x1; x += 32) {
                gl[o + x] = lab;
                gm[o + x] = 1;
            }
        }
        // histogram of the runs over the labels (bucket FUSED_LCAP: all labels beyond the shared table)
        for (int i = tid; i < n_runs; i += T) atomicAdd(&s_hist[min(label_of(P, i) - 1, FUSED_LCAP)], 1);
    }
    __syncthreads();
    if (base < 0) return;

    // ---- 5. per-label accumulators (loki/pipeline.py:589-625) -----------------------------------------
    // Counting sort of the runs by label, then every thread walks a contiguous piece of the sorted list:
    // the sums of the current label stay in registers and are flushed when the label changes.
    if (tid == 0) {
        int acc0 = 0;
        for (int b = 0; b <= FUSED_LCAP; b++) { int c = s_hist[b]; s_hist[b] = acc0; acc0 += c; }
    }
    __syncthreads();
    for (int i = tid; i < n_runs; i += T) rO[atomicAdd(&s_hist[min(label_of(P, i) - 1, FUSED_LCAP)], 1)] = (u16)i;
    __syncthreads();
    const uint8_t *gi = intensity ? intensity + v.pix_off : nullptr;
    const int p_chunk = (n_runs + T - 1) / T;
    const int p_lo = min(tid * p_chunk, n_runs), p_hi = min(p_lo + p_chunk, n_runs);
    {
        int cur = 0;
        u64 aN = 0, aR = 0, aC = 0, aRR = 0, aRC = 0, aCC = 0, aRRR = 0, aRRC = 0, aRCC = 0, aCCC = 0;
        uint32_t aV = 0, aZ = 0;
        int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1;
        uint32_t vmn = FULL, vmx = 0u;
        auto flush = [&](int lab) { // add the register sums of label `lab` to its accumulator row
            u64 *Aa;
            int *Ee;
            const u64 sums[10] = {aN, aR, aC, aRR, aRC, aCC, aRRR, aRRC, aRCC, aCCC};
            if (lab <= FUSED_LCAP) {
                Aa = ACC[lab - 1].a; Ee = ACC[lab - 1].e;
#pragma unroll
                for (int j = 0; j < 10; j++) shared_add64(Aa + j, sums[j]);
                if (gi) { shared_add64(Aa + A_V, (u64)aV); if (aZ) shared_add64(Aa + A_Z, (u64)aZ); }
            } else {
                Aa = acc_stage + (i64)(base + lab - 1) * MAZE_NACC; Ee = ext_stage + (i64)(base + lab - 1) * MAZE_NEXT;
#pragma unroll
                for (int j = 0; j < 10; j++) atomicAdd(Aa + j, sums[j]);
                if (gi) { atomicAdd(Aa + A_V, (u64)aV); atomicAdd(Aa + A_Z, (u64)aZ); }
            }
            atomicMin(Ee + E_RMIN, rmin); atomicMax(Ee + E_RMAX, rmax);
            atomicMin(Ee + E_CMIN, cmin); atomicMax(Ee + E_CMAX, cmax);
            if (gi) {
                uint32_t mn = __vminu4(vmn, vmn >> 16); mn = __vminu4(mn, mn >> 8);
                uint32_t mx = __vmaxu4(vmx, vmx >> 16); mx = __vmaxu4(mx, mx >> 8);
                atomicMin(Ee + E_VMIN, (int)(mn & 0xffu)); atomicMax(Ee + E_VMAX, (int)(mx & 0xffu));
            }
        };
        for (int pos = p_lo; pos < p_hi; pos++) {
            const int i = rO[pos];
            const int L = label_of(P, i);
            if (L != cur) {
                if (cur) flush(cur); // label change inside a chunk: rare (the list is sorted by label)
                aN = aR = aC = aRR = aRC = aCC = aRRR = aRRC = aRCC = aCCC = 0;
                aV = aZ = 0; rmin = 0x7fffffff; rmax = -1; cmin = 0x7fffffff; cmax = -1; vmn = FULL; vmx = 0u;
                cur = L;
            }
            const u64 y = rY[i], a = rX0[i], b = rX1[i], n = b - a + 1;
            const u64 S1 = n * (a + b) / 2;
            const u64 S2 = pw2(b) - (a ? pw2(a - 1) : 0);
            const u64 S3 = pw3(b) - (a ? pw3(a - 1) : 0);
            aN += n; aR += y * n; aC += S1; aRR += y * y * n; aRC += y * S1; aCC += S2;
            aRRR += y * y * y * n; aRRC += y * y * S1; aRCC += y * S2; aCCC += S3;
            rmin = min(rmin, (int)y); rmax = max(rmax, (int)y); cmin = min(cmin, (int)a); cmax = max(cmax, (int)b);
            if (gi) { // intensity of the run's pixels, four at a time (byte-SIMD on aligned words)
                const uint8_t *prow = gi + (size_t)y * W;
                const uint32_t al = (uint32_t)((uintptr_t)prow & 3u);   // row start relative to 4-byte alignment
                const uint32_t *qq = (const uint32_t *)(prow - al);     // word g holds row bytes 4g-al .. 4g-al+3
                const int ga = ((int)a + (int)al) >> 2, gb = ((int)b + (int)al) >> 2;
                for (int g0 = ga; g0 <= gb; g0 += 4) { // four loads in flight: the loop is bound by L2 latency
                    uint32_t pxs[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) pxs[u] = __ldg(qq + min(g0 + u, gb));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int g = g0 + u;
                        const uint32_t px = pxs[u];
                        const int c0 = 4 * g - (int)al;                 // column of byte 0 of this word
                        uint32_t nib = g <= gb ? 0xfu : 0u;             // words past the run contribute nothing
                        if (c0 < (int)a) nib &= 0xfu << ((int)a - c0);
                        if (c0 + 3 > (int)b) nib &= 0xfu >> min(c0 + 3 - (int)b, 31);
                        uint32_t bm = ((nib * 0x00204081u) & 0x01010101u) * 0xffu;
                        aV = __dp4a(px & bm, 0x01010101u, aV);                              // sum of the four masked bytes
                        aZ += __popc(~(((px & 0x7f7f7f7fu) + 0x7f7f7f7fu) | px) & bm & 0x80808080u); // bytes equal to 0
                        vmn = __vminu4(vmn, px | ~bm);
                        vmx = __vmaxu4(vmx, px & bm);
                    }
                }
            }
        }
        // end of chunk: neighbouring threads mostly hold the SAME label -> reduce across the warp, one flush per
        // label and warp (all 32 lanes take part; lanes without work carry label 0)
        uint32_t todo = __ballot_sync(FULL, cur != 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int lab = __shfl_sync(FULL, cur, leader);
            const bool in = (cur == lab);
            todo &= ~__ballot_sync(FULL, in);
            u64 v[10] = {aN, aR, aC, aRR, aRC, aCC, aRRR, aRRC, aRCC, aCCC};
#pragma unroll
            for (int j = 0; j < 10; j++) {
                u64 x = in ? v[j] : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                v[j] = x;
            }
            uint32_t sV = __reduce_add_sync(FULL, in ? aV : 0u), sZ = __reduce_add_sync(FULL, in ? aZ : 0u);
            int r0 = __reduce_min_sync(FULL, in ? rmin : 0x7fffffff), r1 = __reduce_max_sync(FULL, in ? rmax : -1);
            int c0 = __reduce_min_sync(FULL, in ? cmin : 0x7fffffff), c1 = __reduce_max_sync(FULL, in ? cmax : -1);
            uint32_t mn = in ? vmn : FULL, mx = in ? vmx : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                mn = __vminu4(mn, __shfl_xor_sync(FULL, mn, d));
                mx = __vmaxu4(mx, __shfl_xor_sync(FULL, mx, d));
            }
            if (lane == leader) {
                aN = v[0]; aR = v[1]; aC = v[2]; aRR = v[3]; aRC = v[4]; aCC = v[5]; aRRR = v[6]; aRRC = v[7];
                aRCC = v[8]; aCCC = v[9]; aV = sV; aZ = sZ; rmin = r0; rmax = r1; cmin = c0; cmax = c1; vmn = mn; vmx = mx;
                flush(lab);
            }
        }
    }
    __syncthreads();
    if (prm.high_order) {
        // float64 central moments with p + q > 3 about the exact centroid: sum_c (c - cc)^q over a run follows
        // from its n, S1, S2, S3 (columns taken relative to the run start); same walk over the sorted runs
        for (int l = tid; l < n_lab; l += T) {
            const u64 *Aa = l < FUSED_LCAP ? ACC[l].a : acc_stage + (i64)(base + l) * MAZE_NACC;
            double *Hh = l < FUSED_LCAP ? ACC[l].h : hi_stage + (i64)(base + l) * 8;
            double dn = (double)*(const volatile u64 *)(Aa + A_N);
            Hh[H_CR] = (double)*(const volatile u64 *)(Aa + A_R) / dn;
            Hh[H_CC] = (double)*(const volatile u64 *)(Aa + A_C) / dn;
        }
        __syncthreads();
        int cur = 0;
        double h13 = 0, h22 = 0, h31 = 0, h23 = 0, h32 = 0, h33 = 0, cr = 0, cc = 0;
        auto hrow = [&](int lab) { return lab <= FUSED_LCAP ? ACC[lab - 1].h : hi_stage + (i64)(base + lab - 1) * 8; };
        for (int pos = p_lo; pos < p_hi; pos++) {
            const int i = rO[pos];
            const int L = label_of(P, i);
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
            // columns relative to the run start: j = 0 .. n-1, u = j - o with o = cc - x0
            const int nn = (int)rX1[i] - (int)rX0[i] + 1;
            const double dn = (double)nn, m1 = (double)(nn - 1);
            const double S1 = dn * m1 * 0.5;                          // sum j
            const double S2 = m1 * dn * (2.0 * m1 + 1.0) / 6.0;       // sum j^2
            const double S3 = S1 * S1;                                // sum j^3
            const double o = cc - (double)rX0[i];
            const double T1 = S1 - dn * o;
            const double T2 = S2 - 2.0 * o * S1 + dn * o * o;
            const double T3 = S3 - 3.0 * o * S2 + 3.0 * o * o * S1 - dn * o * o * o;
            const double dr = (double)rY[i] - cr, dr2 = dr * dr, dr3 = dr2 * dr;
            h13 += dr * T3; h22 += dr2 * T2; h31 += dr3 * T1;
            h23 += dr2 * T3; h32 += dr3 * T2; h33 += dr3 * T3;
        }
        uint32_t todo = __ballot_sync(FULL, cur != 0);
        while (todo) { // one flush per label and warp
            const int leader = __ffs(todo) - 1;
            const int lab = __shfl_sync(FULL, cur, leader);
            const bool in = (cur == lab);
            todo &= ~__ballot_sync(FULL, in);
            double v[6] = {h13, h22, h31, h23, h32, h33};
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double x = in ? v[j] : 0.0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                v[j] = x;
            }
            if (lane == leader) {
                double *Hh = hrow(lab);
                atomicAdd(Hh + H_13, v[0]); atomicAdd(Hh + H_22, v[1]); atomicAdd(Hh + H_31, v[2]);
                atomicAdd(Hh + H_23, v[3]); atomicAdd(Hh + H_32, v[4]); atomicAdd(Hh + H_33, v[5]);
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
        prm.pass[p].use_phantom = 1;
        for (int i = 0; i <= MAZE_MAX_DISK_RADIUS; i++) prm.pass[p].w[i] = 0;
    }
    for (int p = 0; p < n_pass; p++) {
        prm.pass[p].invert = pass_invert_host[p] ? 1 : 0;
        if (!maze_pass_table(pass_t_host[p], &prm.pass[p].R, prm.pass[p].w, &prm.pass[p].use_phantom)) return MAZE_ERR_BADARG;
    }
    MAZE_CUDA(cudaMemsetAsync(stage_counter, 0, sizeof(int32_t), s), "stage counter");
    FusedArgs a = {image, intensity, vig, bits, mask, labels, n_labels, fallback, acc_base, stage_counter,
                   (u64 *)acc_stage, hi_stage, ext_stage};
    const int caps[MAZE_FUSED_CLASSES] = MAZE_FUSED_CAPS;
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
        // threads per CTA by class: finer shared-memory classes let CTAs of different sizes share an SM
        if (caps[c] <= 1024) rc = launch_class<128>(n, caps[c], sc, list, prm, a);
        else if (caps[c] <= 4096) rc = launch_class<256>(n, caps[c], sc, list, prm, a);
        else if (caps[c] <= 19456) rc = launch_class<512>(n, caps[c], sc, list, prm, a);
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
    static thread_local int attr_dev = -1;
    int dev = 0;
    MAZE_CUDA(cudaGetDevice(&dev), "get device");
    if (attr_dev != dev) { // same carve-out as the stage kernels, so that it can share an SM with them
        MAZE_CUDA(cudaFuncSetAttribute(k_count_scan, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared), "scan carveout");
        attr_dev = dev;
    }
    MAZE_KERNEL(KID_COUNT_SCAN, (cudaStream_t)stream,
                k_count_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(n_labels, n_img, lab_off));
    return MAZE_OK;
}
