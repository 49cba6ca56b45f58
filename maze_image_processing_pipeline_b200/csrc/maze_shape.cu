// Per-label shape features that need the pixel neighbourhood: the numbers
// skimage's RegionProperties hands to morphocut's CalculateZooProcessFeatures (loki/pipeline.py:625, 654) as
// `perimeter`, `euler_number` and `filled_area`.
//
//   perimeter     skimage.measure.perimeter(region.image, neighborhood=4): border = image minus its erosion by
//                 the 4-neighbourhood cross (outside the crop counts as background); every border pixel gets the
//                 code 1 + 2 * (#4-neighbours on the border) + 10 * (#diagonal neighbours on the border); codes
//                 5, 7, 15, 17, 25, 27 weigh 1, codes 21, 33 weigh sqrt(2), codes 13, 23 weigh (1 + sqrt(2)) / 2.
//   euler_number  skimage.measure.euler_number(region.image, connectivity=2): over all 2x2 windows of the
//                 zero-padded crop, +1 for "only the top-left pixel set", -1 for "top-right and bottom-left set,
//                 the other two clear", -1 for "all but the bottom-right set".
//   convex_area   np.sum(skimage.morphology.convex_hull_image(region.image)): hull of the four edge midpoints of every
//                 pixel, pixel centres inside or on the hull count.
//   filled_area   region.image with its holes filled by scipy.ndimage.binary_fill_holes(image, ones((3, 3))): the
//                 complement is flooded from outside the crop with 8-CONNECTED steps; what the flood does not
//                 reach is object or hole.
//
// One CTA takes one object at a time from a work counter: it builds the object's own bit plane P (label == l
// inside the bounding box, one zero pixel of frame around it) from the label image, derives the border plane,
// the three perimeter class counts and the Euler sum with 32-pixel word operations, then floods the complement
// from the frame: one row-parallel pass (in-word Kogge-Stone fill, carries across the words of a row resolved
// with one ballot + add per direction), then sweeps down and up that only touch rows whose neighbours bring new
// seeds, until a whole round changes nothing (a convex object is done after the first pass).  Planes of small crops live in shared
// memory, the others in the CTA's slab of a caller-provided pool.  The convex hull area follows from the row
// extremes of the plane (monotone chain + exact rasterisation).  All counts are exact integers; the only
// floating-point step is the final weighted sum of the perimeter.
#include "maze_common.cuh"

#define SH_T 128           /* threads per CTA, crops whose planes fit shared memory */
#define SH_T_BIG 512       /* threads per CTA for the others (planes in the CTA's slab): 16 bands of rows flood at once */
#define SH_SMEM_WORDS 14000 /* both planes of a crop of up to 7000 words (~470 x 470 px) stay in shared memory:
                               56 KB per CTA, four CTAs per SM; a row step on shared planes costs tens of cycles, on
                               the global slab an L2 round trip */
#define SH_MAX_CHUNKS 5    /* 32-word chunks per framed row: (4096 + 2 + 31) / 32 = 129 words */

// seeds spread along the runs of m, both directions, inside one word
__device__ __forceinline__ uint32_t fill_in_word(uint32_t seeds, uint32_t m)
{
    uint32_t s = seeds & m, p = m;
    s |= p & (s << 1); p &= p << 1;
    s |= p & (s << 2); p &= p << 2;
    s |= p & (s << 4); p &= p << 4;
    s |= p & (s << 8); p &= p << 8;
    s |= p & (s << 16);
    p = m;
    s |= p & (s >> 1); p &= p >> 1;
    s |= p & (s >> 2); p &= p >> 2;
    s |= p & (s >> 4); p &= p >> 4;
    s |= p & (s >> 8); p &= p >> 8;
    s |= p & (s >> 16);
    return s;
}

__device__ __forceinline__ uint32_t trailing_ones(uint32_t m)
{
    int tz = __ffs(~m) - 1;  // position of the lowest zero; -1 when m is all ones
    return tz < 0 ? FULL : ((1u << tz) - 1u);
}

__device__ __forceinline__ uint32_t leading_ones(uint32_t m)
{
    int lz = __clz(~m);  // number of leading ones of m (32 when m is all ones)
    return lz ? (FULL << (32 - lz)) : 0u;
}

__device__ __forceinline__ uint32_t ldv(const uint32_t *p) { return *(const volatile uint32_t *)p; }

// bit-sliced count of four one-bit planes: c0, c1, c2 = bits of x1 + x2 + x3 + x4
__device__ __forceinline__ void count4(uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t &c0, uint32_t &c1,
                                       uint32_t &c2)
{
    uint32_t p1 = x1 ^ x2, h1 = x1 & x2, p2 = x3 ^ x4, h2 = x3 & x4;
    c0 = p1 ^ p2;
    uint32_t carry = p1 & p2;
    c1 = h1 ^ h2 ^ carry;
    c2 = h1 & h2;
}

template <int T, bool BIG>
__global__ void __launch_bounds__(T) k_label_shape(const int32_t *__restrict__ labels,
                                                      const uint32_t *__restrict__ bits,
                                                      const maze_vignette_t *__restrict__ vig,
                                                      const double *__restrict__ table, int n_obj, uint32_t *pool,
                                                      i64 slab_words, int big_chain_words, int runs, int *work_counter,
                                                      double *__restrict__ shape)
{
    extern __shared__ uint32_t s_planes[];
    __shared__ int s_job;
    __shared__ int s_acc[6];  // n1, n2, n3, euler, reached, convex
    __shared__ int s_nchain[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T / 32;
    uint32_t *slab = pool + (i64)blockIdx.x * slab_words;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_job = atomicAdd(work_counter, 1);
        if (tid < 6) s_acc[tid] = 0;
        __syncthreads();
        const int o = s_job;
        if (o >= n_obj) return;
        const double *row = table + (i64)o * MAZE_NFEAT;
        double *out = shape + (i64)o * MAZE_NSHAPE;
        const double area = row[MAZE_F_AREA];
        if (!(area > 0)) {  // label removed by a filter, or absent
            if (!BIG && tid < MAZE_NSHAPE) out[tid] = nan("");
            continue;
        }
        const int r0 = (int)row[MAZE_F_BBOX], c0 = (int)row[MAZE_F_BBOX + 1];
        const int h = (int)row[MAZE_F_BBOX + 2] - r0, w = (int)row[MAZE_F_BBOX + 3] - c0;
        const int label = (int)row[MAZE_F_LABEL];
        const maze_vignette_t v = vig[(int)row[MAZE_F_IMAGE]];
        const int rows = h + 2, fw = w + 2, cw = (fw + 31) >> 5, nwords = rows * cw;
        if ((2 * nwords + 4 * h + 2 > SH_SMEM_WORDS) != BIG) continue;  // the other launch takes this object
        if (BIG && (i64)2 * nwords > slab_words) {          // cannot happen with a slab sized by the caller
            if (tid < MAZE_NSHAPE) out[tid] = nan("");
            continue;
        }
        uint32_t *P = BIG ? slab : s_planes;
        uint32_t *R = P + nwords;

        // ---- A. the object's own plane, framed by one zero pixel --------------------------------------------
        // work items of (row, group of eight words); four items = 32 loads per lane are in flight, the loop is bound
        // by memory latency
        if (runs) {
            // labels are constant along the runs of `bits` (output of maze_label / the fused kernel, also after the
            // label filters): one thread per word takes the 32 bits of the bit plane and keeps the runs whose first
            // pixel carries this label -- one or two 4-byte loads per word instead of 32
            for (int i = tid; i < nwords; i += T) {
                const int fy = i / cw, k = i - fy * cw;
                uint32_t word = 0;
                if (fy >= 1 && fy <= h) {
                    const int y = r0 + fy - 1;
                    const int x0 = c0 - 1 + 32 * k;                       // pixel column of bit 0 (may be -1)
                    const uint32_t *brow = bits + v.word_off + (i64)y * v.wpr;
                    const int wi = x0 >> 5, sh = x0 & 31;                 // arithmetic shift: x0 = -1 -> wi = -1, sh = 31
                    const uint32_t lo = (wi >= 0 && wi < v.wpr) ? __ldg(brow + wi) : 0u;
                    const uint32_t hi = (wi + 1 >= 0 && wi + 1 < v.wpr) ? __ldg(brow + wi + 1) : 0u;
                    uint32_t m = __funnelshift_r(lo, hi, sh);
                    // crop: framed columns 1 .. w
                    const int f_lo = max(1 - 32 * k, 0), f_hi = min(w - 32 * k, 31);  // bit range inside this word
                    if (f_hi < f_lo) m = 0;
                    else m &= (f_hi == 31 ? FULL : ((2u << f_hi) - 1u)) & ~((1u << f_lo) - 1u);
                    const int32_t *lrow = labels + v.pix_off + (i64)y * v.w;
                    uint32_t rest = m;
                    while (rest) {
                        const int b = __ffs(rest) - 1;                     // first pixel of the next run in the word
                        const uint32_t from_b = rest >> b;                 // run = the ones from bit b up to the next zero
                        const int len = (~from_b == 0u) ? 32 - b : __ffs(~from_b) - 1;
                        const uint32_t run = (len >= 32 ? FULL : ((1u << len) - 1u)) << b;
                        if (__ldg(lrow + x0 + b) == label) word |= run;
                        rest &= ~run;
                    }
                }
                P[i] = word;
            }
        } else {
            const int ngrp = (cw + 7) >> 3, nitem = rows * ngrp;
            for (int it0 = warp * 4; it0 < nitem; it0 += nwarp * 4) {
                uint32_t cmp[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int it = it0 + q;
                    const int fy = it / ngrp, k0 = (it - fy * ngrp) * 8;
                    const bool inside = it < nitem && fy >= 1 && fy <= h;
                    const int y = r0 + fy - 1;
                    cmp[q] = 0;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int fx = 32 * (k0 + u) + lane, x = c0 - 1 + fx;
                        bool b = false;
                        if (inside && fx >= 1 && fx <= w) {
                            if (labels) b = __ldg(labels + v.pix_off + (i64)y * v.w + x) == label;
                            else b = (__ldg(bits + v.word_off + (i64)y * v.wpr + (x >> 5)) >> (x & 31)) & 1u;
                        }
                        cmp[q] |= (b ? 1u : 0u) << u;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int it = it0 + q;
                    const int fy = it / ngrp, k0 = (it - fy * ngrp) * 8;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const uint32_t word = __ballot_sync(FULL, (cmp[q] >> u) & 1u);
                        if (lane == u && it < nitem && k0 + u < cw) P[fy * cw + k0 + u] = word;
                    }
                }
            }
        }
        __syncthreads();

        // ---- B. border plane (into R) and the Euler sum ------------------------------------------------------
        int euler = 0;
        for (int i = tid; i < nwords; i += T) {
            const int fy = i / cw, k = i - fy * cw;
            const uint32_t C = P[i];
            const uint32_t U = fy > 0 ? P[i - cw] : 0u, D = fy < rows - 1 ? P[i + cw] : 0u;
            const uint32_t Lw = k > 0 ? P[i - 1] : 0u, Rw = k < cw - 1 ? P[i + 1] : 0u;
            const uint32_t ULw = (fy > 0 && k > 0) ? P[i - cw - 1] : 0u;
            const uint32_t Cl = (C << 1) | (Lw >> 31), Cr = (C >> 1) | (Rw << 31);  // pixel x-1 / x+1 at bit x
            R[i] = C & ~(U & D & Cl & Cr);
            // window a = (y-1, x-1), b = (y-1, x), c = (y, x-1), d = (y, x)
            const uint32_t Ul = (U << 1) | (ULw >> 31);
            euler += __popc(Ul & ~U & ~Cl & ~C) - __popc(~Ul & U & Cl & ~C) - __popc(Ul & U & Cl & ~C);
        }
        __syncthreads();

        // ---- C. perimeter classes ----------------------------------------------------------------------------
        int n1 = 0, n2 = 0, n3 = 0;
        for (int i = tid; i < nwords; i += T) {
            const uint32_t B = R[i];
            if (!B) continue;  // border pixels exist only in rows 1..h
            const int fy = i / cw, k = i - fy * cw;
            const bool hl = k > 0, hr = k < cw - 1;
            const uint32_t Bu = R[i - cw], Bd = R[i + cw];
            const uint32_t Bl = hl ? R[i - 1] : 0u, Br = hr ? R[i + 1] : 0u;
            const uint32_t Bul = hl ? R[i - cw - 1] : 0u, Bur = hr ? R[i - cw + 1] : 0u;
            const uint32_t Bdl = hl ? R[i + cw - 1] : 0u, Bdr = hr ? R[i + cw + 1] : 0u;
            uint32_t a0, a1, a2, b0, b1, b2;
            count4(Bu, Bd, (B << 1) | (Bl >> 31), (B >> 1) | (Br << 31), a0, a1, a2);
            count4((Bu << 1) | (Bul >> 31), (Bu >> 1) | (Bur << 31), (Bd << 1) | (Bdl >> 31), (Bd >> 1) | (Bdr << 31), b0,
                   b1, b2);
            const uint32_t a_is0 = ~(a0 | a1 | a2), a_is1 = a0 & ~a1, a_is23 = a1;  // a1 set <=> count 2 or 3
            const uint32_t b_le2 = ~((b1 & b0) | b2), b_is1 = b0 & ~b1, b_is2 = b1 & ~b0, b_is3 = b1 & b0;
            n1 += __popc(B & a_is23 & b_le2);                           // codes 5, 7, 15, 17, 25, 27
            n2 += __popc(B & ((a_is0 & b_is2) | (a_is1 & b_is3)));      // codes 21, 33
            n3 += __popc(B & a_is1 & (b_is1 | b_is2));                  // codes 13, 23
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            n1 += __shfl_xor_sync(FULL, n1, d);
            n2 += __shfl_xor_sync(FULL, n2, d);
            n3 += __shfl_xor_sync(FULL, n3, d);
            euler += __shfl_xor_sync(FULL, euler, d);
        }
        if (lane == 0) {
            atomicAdd(&s_acc[0], n1);
            atomicAdd(&s_acc[1], n2);
            atomicAdd(&s_acc[2], n3);
            atomicAdd(&s_acc[3], euler);
        }
        __syncthreads();

        // ---- D. flood of the complement from the frame (8-connected) ------------------------------------------
        for (int i = tid; i < nwords; i += T) {
            const int fy = i / cw, k = i - fy * cw;
            uint32_t s = 0;
            if (fy == 0 || fy == rows - 1) s = valid_mask(fw, k);
            else {
                if (k == 0) s |= 1u;
                if (k == (fw - 1) >> 5) s |= 1u << ((fw - 1) & 31);
            }
            R[i] = s;
        }
        __syncthreads();
        // one row: seeds = what the row already holds plus the 8-neighbour spread of the rows above and below;
        // in-word fill, then the carries across the words of the row.  `lazy` skips rows the neighbours add nothing
        // to (every row is closed under its own fill after the first, row-parallel pass).  Returns "row changed".
        const int nchunk = (cw + 31) >> 5;
        auto flood_row = [&](int fy, bool lazy) -> int {
            uint32_t s[SH_MAX_CHUNKS], M[SH_MAX_CHUNKS], old[SH_MAX_CHUNKS];
            bool fresh = false;
#pragma unroll
            for (int c = 0; c < SH_MAX_CHUNKS; c++) {
                if (c < nchunk) {
                    const int k = 32 * c + lane;
                    const bool ok = k < cw;
                    const int i = fy * cw + k;
                    M[c] = ok ? (~P[i] & valid_mask(fw, k)) : 0u;
                    old[c] = ok ? ldv(R + i) : 0u;
                    uint32_t V = 0, VL = 0, VR = 0;
                    if (ok) {
                        V = ldv(R + i - cw) | ldv(R + i + cw);
                        if (k > 0) VL = ldv(R + i - cw - 1) | ldv(R + i + cw - 1);
                        if (k < cw - 1) VR = ldv(R + i - cw + 1) | ldv(R + i + cw + 1);
                    }
                    const uint32_t spread = (V | (V << 1) | (VL >> 31) | (V >> 1) | (VR << 31)) & M[c];
                    fresh |= (spread & ~old[c]) != 0u;
                    s[c] = old[c] | spread;
                }
            }
            if (lazy && !__any_sync(FULL, fresh)) return 0;
            uint32_t cin = 0;
#pragma unroll
            for (int c = 0; c < SH_MAX_CHUNKS; c++) {  // local fill, then carries towards larger x
                if (c < nchunk) {
                    s[c] = fill_in_word(s[c], M[c]);
                    const uint32_t G = __ballot_sync(FULL, s[c] >> 31), Pm = __ballot_sync(FULL, M[c] == FULL);
                    const uint32_t A = G | Pm;
                    const u64 S = (u64)A + G + cin;
                    const uint32_t Cin = (uint32_t)S ^ A ^ G;
                    cin = (uint32_t)(S >> 32);
                    if ((Cin >> lane) & 1u) s[c] |= trailing_ones(M[c]);
                }
            }
            cin = 0;
            int changed = 0;
#pragma unroll
            for (int c = SH_MAX_CHUNKS - 1; c >= 0; c--) {  // carries towards smaller x
                if (c < nchunk) {
                    const uint32_t G = __brev(__ballot_sync(FULL, s[c] & 1u));
                    const uint32_t Pm = __brev(__ballot_sync(FULL, M[c] == FULL));
                    const uint32_t A = G | Pm;
                    const u64 S = (u64)A + G + cin;
                    const uint32_t Cin = (uint32_t)S ^ A ^ G;
                    cin = (uint32_t)(S >> 32);
                    if ((Cin >> (31 - lane)) & 1u) s[c] |= leading_ones(M[c]);
                    const int k = 32 * c + lane;
                    if (k < cw && s[c] != old[c]) {
                        R[fy * cw + k] = s[c];
                        changed = 1;
                    }
                }
            }
            __syncwarp();
            return changed;
        };
        // first pass, all rows in parallel: from the frame columns (and whatever the neighbours already hold)
        for (int fy = 1 + warp; fy < rows - 1; fy += nwarp) flood_row(fy, false);
        __syncthreads();
        // then sweeps down and up over bands of rows (one band per warp) until a whole round changes nothing;
        // for a convex object the first pass has reached everything and the sweeps only look
        const int band = (rows + nwarp - 1) / nwarp;
        const int b_lo = max(1, warp * band), b_hi = min(rows - 1, (warp + 1) * band);  // frame rows are complete
        for (;;) {
            int changed = 0;
            for (int fy = b_lo; fy < b_hi; fy++) changed |= flood_row(fy, true);
            for (int fy = b_hi - 1; fy >= b_lo; fy--) changed |= flood_row(fy, true);
            if (!__syncthreads_or(changed)) break;
        }
        int reached = 0;
        for (int i = tid; i < nwords; i += T) reached += __popc(R[i]);
#pragma unroll
        for (int d = 16; d; d >>= 1) reached += __shfl_xor_sync(FULL, reached, d);
        if (lane == 0) atomicAdd(&s_acc[4], reached);
        __syncthreads();

        // ---- E. area of the convex hull image (skimage.morphology.convex_hull_image: every pixel contributes the
        // midpoints of its four edges; a pixel is in when its centre lies in or on the hull) --------------------------
        // Doubled coordinates (row y <-> Y = 2y): per level Y = -1 .. 2h-1 the leftmost / rightmost hull candidate
        // follows from the row extremes; two threads run the monotone chain in place; every row then counts the
        // pixel centres between the two chains in exact integer arithmetic.
        uint32_t *CL = BIG ? s_planes : s_planes + 2 * nwords, *CR = CL + (2 * h + 1);
        const bool hull_ok = BIG ? (4 * h + 2 <= big_chain_words) : true;
        if (hull_ok) {
            for (int y = tid; y < h; y += T) {  // row extremes (pixel columns) into the now idle plane R
                const uint32_t *prow = P + (y + 1) * cw;
                int xl = -1, xr = -1;
                for (int k = 0; k < cw; k++) {
                    const uint32_t m = prow[k];
                    if (m) { xl = 32 * k + __ffs(m) - 2; break; }
                }
                for (int k = cw - 1; k >= 0 && xl >= 0; k--) {
                    const uint32_t m = prow[k];
                    if (m) { xr = 32 * k + 30 - __clz(m); break; }
                }
                R[y] = xl < 0 ? FULL : (((uint32_t)xl << 16) | (uint32_t)xr);
            }
            __syncthreads();
            for (int j = tid; j <= 2 * h; j += T) {  // level Y = j - 1, stored as (Y + 1) << 16 | (X + 1)
                int l = 0x7fffffff, r = -0x7fffffff;
                if (j & 1) {  // pixel row: (y, xl - 1/2), (y, xr + 1/2)
                    const uint32_t e = R[j >> 1];
                    if (e != FULL) { l = 2 * (int)(e >> 16) - 1; r = 2 * (int)(e & 0xffffu) + 1; }
                } else {      // between rows: (y + 1/2, x) of the row above, (y - 1/2, x) of the row below
                    const int ya = (j >> 1) - 1, yb = j >> 1;
                    if (ya >= 0) {
                        const uint32_t e = R[ya];
                        if (e != FULL) { l = min(l, 2 * (int)(e >> 16)); r = max(r, 2 * (int)(e & 0xffffu)); }
                    }
                    if (yb < h) {
                        const uint32_t e = R[yb];
                        if (e != FULL) { l = min(l, 2 * (int)(e >> 16)); r = max(r, 2 * (int)(e & 0xffffu)); }
                    }
                }
                const bool has = r >= l;
                CL[j] = has ? (((uint32_t)j << 16) | (uint32_t)(l + 1)) : FULL;
                CR[j] = has ? (((uint32_t)j << 16) | (uint32_t)(r + 1)) : FULL;
            }
            __syncthreads();
            if (tid == 0 || tid == 32) {  // monotone chain, in place (the stack never passes the read position)
                uint32_t *C = tid == 0 ? CL : CR;
                const bool left = tid == 0;
                int n = 0;
                for (int j = 0; j <= 2 * h; j++) {
                    const uint32_t c = C[j];
                    if (c == FULL) continue;
                    const int cy = (int)(c >> 16), cx = (int)(c & 0xffffu);
                    while (n >= 2) {
                        const uint32_t a = C[n - 2], b = C[n - 1];
                        const int ay = (int)(a >> 16), ax = (int)(a & 0xffffu), by = (int)(b >> 16), bx = (int)(b & 0xffffu);
                        const int cr = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay);
                        if (left ? cr >= 0 : cr <= 0) n--; else break;
                    }
                    C[n++] = c;
                }
                s_nchain[left ? 0 : 1] = n;
            }
            __syncthreads();
            int cnt = 0;
            const int nl = s_nchain[0], nr = s_nchain[1];
            for (int y = tid; y < h; y += T) {
                const int Yp = 2 * y + 1;  // stored level of the row
                if (nl < 2 || nr < 2 || Yp < (int)(CL[0] >> 16) || Yp > (int)(CL[nl - 1] >> 16)) continue;
                int bound[2];
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const uint32_t *C = side ? CR : CL;
                    const int n = side ? nr : nl;
                    int lo = 0, hi = n - 2;  // last segment start with level <= Yp
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if ((int)(C[mid] >> 16) <= Yp) lo = mid; else hi = mid - 1;
                    }
                    const uint32_t a = C[lo], b = C[lo + 1];
                    const int y1 = (int)(a >> 16), x1 = (int)(a & 0xffffu) - 1, y2 = (int)(b >> 16), x2 = (int)(b & 0xffffu) - 1;
                    const int D = y2 - y1, N = x1 * D + (x2 - x1) * (Yp - y1);  // hull abscissa (doubled) = N / D
                    const int q = 2 * D;
                    // left: smallest x with 2x >= N / D; right: largest x with 2x <= N / D
                    if (side == 0) bound[0] = N >= 0 ? (N + q - 1) / q : -((-N) / q);
                    else bound[1] = N >= 0 ? N / q : -((-N + q - 1) / q);
                }
                const int xa = max(bound[0], 0), xb = min(bound[1], w - 1);
                cnt += max(0, xb - xa + 1);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
            if (lane == 0) atomicAdd(&s_acc[5], cnt);
            __syncthreads();
        }
        if (tid == 0) {
            const double SQ2 = 1.4142135623730951;
            const int m1 = s_acc[0], m2 = s_acc[1], m3 = s_acc[2];
            out[MAZE_S_PERIMETER] = (double)m1 + (double)m2 * SQ2 + (double)m3 * ((1.0 + SQ2) / 2.0);
            out[MAZE_S_FILLED_AREA] = (double)((i64)rows * fw - s_acc[4]);
            out[MAZE_S_EULER] = (double)s_acc[3];
            out[MAZE_S_N1] = (double)m1;
            out[MAZE_S_N2] = (double)m2;
            out[MAZE_S_N3] = (double)m3;
            out[MAZE_S_CONVEX_AREA] = hull_ok ? (double)s_acc[5] : nan("");
            out[7] = nan("");
        }
    }
}

struct ShapeFork {  // side stream for the launch of the large crops (per host thread and device)
    int device;
    cudaStream_t aux;
    cudaEvent_t fork, join;
    int big_smem;  // dynamic shared memory the large-crop kernel is configured for
};

static ShapeFork *shape_fork()
{
    static thread_local ShapeFork pool[16];
    static thread_local int n_pool = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < n_pool; i++)
        if (pool[i].device == dev) return &pool[i];
    if (n_pool >= 16) return nullptr;
    ShapeFork *f = &pool[n_pool];
    f->device = dev;
    f->big_smem = 48 * 1024;
    if (cudaStreamCreateWithFlags(&f->aux, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f->join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaFuncSetAttribute(k_label_shape<SH_T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SH_SMEM_WORDS * (int)sizeof(uint32_t)) != cudaSuccess)
        return nullptr;
    n_pool++;
    return f;
}

extern "C" int maze_label_shape(const int32_t *labels, const uint32_t *bits, const maze_vignette_t *vig,
                                const double *table, int n_obj, uint32_t *pool, long long slab_words, int n_slabs,
                                int max_h, int runs, int32_t *work_counter, double *shape, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (n_obj <= 0) return MAZE_OK;
    // every object needs a home: the ones whose planes exceed shared memory live in the slabs
    if ((!labels && !bits) || !vig || !table || !shape || !work_counter || n_slabs <= 0 || slab_words <= 0 || !pool)
        return MAZE_ERR_BADARG;
    if (runs && (!labels || !bits)) return MAZE_ERR_BADARG;
    MAZE_CUDA(cudaMemsetAsync(work_counter, 0, 2 * sizeof(int32_t), s), "label_shape counter");
    ShapeFork *fk = shape_fork();
    if (!fk) return MAZE_ERR_CUDA;
    // the few objects whose planes do not fit shared memory (they take longest: 512 threads each on a slab of the
    // pool) run on a side stream next to everything else (four CTAs of 128 threads per SM, planes in shared memory)
    {
        MAZE_CUDA(cudaEventRecord(fk->fork, s), "label_shape fork");
        MAZE_CUDA(cudaStreamWaitEvent(fk->aux, fk->fork, 0), "label_shape fork wait");
        const int grid_big = n_slabs < n_obj ? n_slabs : n_obj;
        // chain storage of the convex hull step: 2 * (2 * max_h + 1) words of shared memory (objects taller than
        // what 200 KB hold get NaN for the convex area)
        int chain_words = 4 * (max_h > 0 ? max_h : 0) + 8;
        if (chain_words > 50000) chain_words = 50000;
        if (chain_words * (int)sizeof(uint32_t) > fk->big_smem) {
            MAZE_CUDA(cudaFuncSetAttribute(k_label_shape<SH_T_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           chain_words * (int)sizeof(uint32_t)), "label_shape chain smem");
            fk->big_smem = chain_words * (int)sizeof(uint32_t);
        }
        MAZE_KERNEL(KID_LABEL_SHAPE, fk->aux,
                    (k_label_shape<SH_T_BIG, true><<<grid_big, SH_T_BIG, chain_words * sizeof(uint32_t), fk->aux>>>(
                        labels, bits, vig, table, n_obj, pool, slab_words, chain_words, runs ? 1 : 0, work_counter, shape)));
        MAZE_CUDA(cudaEventRecord(fk->join, fk->aux), "label_shape join");
    }
    const int grid = n_obj < 148 * 8 ? n_obj : 148 * 8;
    MAZE_KERNEL(KID_LABEL_SHAPE, s,
                (k_label_shape<SH_T, false><<<grid, SH_T, SH_SMEM_WORDS * sizeof(uint32_t), s>>>(
                    labels, bits, vig, table, n_obj, pool, slab_words, 0, runs ? 1 : 0, work_counter + 1, shape)));
    MAZE_CUDA(cudaStreamWaitEvent(s, fk->join, 0), "label_shape join wait");
    return MAZE_OK;
}
