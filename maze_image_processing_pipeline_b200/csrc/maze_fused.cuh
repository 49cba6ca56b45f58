// Device helpers shared by the vignette-resident kernel (maze_fused.cu) and the band pipeline (maze_bands.cu):
// accumulator layout, block scan, 16-bit union-find, column-walk morphology, SWAR threshold.
#pragma once
#include "maze_common.cuh"

#define FUSED_LCAP 24 /* labels per vignette whose accumulators live in shared memory */

enum { A_N = 0, A_R, A_C, A_RR, A_RC, A_CC, A_RRR, A_RRC, A_RCC, A_CCC, A_V, A_Z };
enum { E_RMIN = 0, E_RMAX, E_CMIN, E_CMAX, E_VMIN, E_VMAX };
enum { H_13 = 0, H_22, H_31, H_23, H_32, H_33, H_CR, H_CC };

struct FusedPass {
    int R;
    int invert;
    int use_phantom; // EDT pass (scipy's phantom pixel applies) or plain binary morphology with a registered footprint
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

// 64-bit add to a shared-memory slot as one or two NATIVE 32-bit atomics (a 64-bit shared atomicAdd is a
// compare-and-swap loop, and all warps of the CTA flush into the same few rows at the same moment): the low half
// accumulates modulo 2^32, the carry of each add follows from the value the atomic returns.
__device__ __forceinline__ void shared_add64(u64 *slot, u64 v)
{
    uint32_t *w = (uint32_t *)slot;
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const uint32_t old = atomicAdd(w, lo);
    const uint32_t up = hi + ((old + lo) < old ? 1u : 0u);
    if (up) atomicAdd(w + 1, up);
}

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

__device__ __forceinline__ u64 pw2(u64 m) { return m * (m + 1) * (2 * m + 1) / 6; }   // sum_{i<=m} i^2, m < 2^16
__device__ __forceinline__ u64 pw3(u64 m) { u64 t = m * (m + 1) / 2; return t * t; }      // sum_{i<=m} i^3
// power sums of 0..m for m < 1024 (columns relative to a 1024-pixel chunk): 32-bit arithmetic suffices
__device__ __forceinline__ uint32_t f_pow2sum(uint32_t m) { return m * (m + 1) / 2 * (2 * m + 1) / 3; }
__device__ __forceinline__ u64 f_pow3sum(uint32_t m) { uint32_t t = m * (m + 1) / 2; return (u64)t * t; }

// One thresholded-EDT pass as a COLUMN WALK: a thread owns word column k and a strip of rows and slides
// down, keeping for the last 2R+1 rows the chord tests at the R+1 chord widths in registers, so every
// word is loaded once and every chord test is computed once (instead of 2R+1 times).
struct WRuntime { // chord half-widths read from the pass table
    const int *p;
    __device__ __forceinline__ int w(int j) const { return p[j]; }
};
template <int A, int B, int C, int D>
struct WFixed { // compile-time chord half-widths of the common small disks: the chord loops unroll completely
    __device__ __forceinline__ constexpr int w(int j) const { return j == 0 ? A : j == 1 ? B : j == 2 ? C : D; }
};

template <int R, int T, typename WT>
__device__ __forceinline__ void morph_columns(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int H, int W,
                                              int wpr, const WT wt, uint32_t inv, bool phantom, int nstrip, int S,
                                              uint32_t inv_next, bool &hz_next, int hz_lo = 0, int hz_hi = 0x7fffffff)
{   // hz_lo / hz_hi: rows whose output takes part in the "has a 0" flag (a band only answers for its own rows)
    int wd[R + 1];
#pragma unroll
    for (int j = 0; j <= R; j++) wd[j] = wt.w(j);
    const int nitems = wpr * nstrip;
    uint32_t hzacc = 0;  // valid bits of the output that are 0 after the NEXT pass's inversion (phantom test)
    for (int q = threadIdx.x; q < nitems; q += T) {
        const int s = q / wpr, k = q - s * wpr;
        const int y0 = s * S, y1 = min(H, y0 + S);
        if (y0 >= y1) continue;
        // neighbour words without branches: a missing neighbour is read from the word itself and forced to ones
        const bool hasL = k > 0, hasR = k + 1 < wpr;
        const uint32_t padC = ~valid_mask(W, k);
        const uint32_t mL = hasL ? 0u : FULL, mR = hasR ? ~valid_mask(W, k + 1) : FULL;
        const int oL = hasL ? -1 : 0, oR = hasR ? 1 : 0;
        uint32_t win[2 * R + 1][R + 1]; // win[i][j]: row (y - R + i) of the current output row y, chord width wd[j]
#pragma unroll
        for (int i = 0; i < 2 * R + 1; i++)
#pragma unroll
            for (int j = 0; j <= R; j++) win[i][j] = FULL;
        auto feed = [&](uint32_t C, uint32_t L, uint32_t Rw) {
            // slide the window up by one row, then the chord tests of the new row, narrow to wide
#pragma unroll
            for (int i = 0; i < 2 * R; i++)
#pragma unroll
                for (int j = 0; j <= R; j++) win[i][j] = win[i + 1][j];
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
        };
        auto emit = [&](uint32_t *out, int orow) {
            uint32_t acc = FULL;
#pragma unroll
            for (int i = 0; i < 2 * R + 1; i++) acc &= win[i][i < R ? R - i : i - R];
            const uint32_t o = (acc ^ inv) & ~padC;
            *out = o;
            if (orow >= hz_lo && orow < hz_hi) hzacc |= ~((o ^ inv_next) | padC);
        };
        // rows fed: [y0 - R, y1 + R); rows outside the image are all ones (the border is not background), except
        // scipy's phantom pixel at (-1, 0) of a plane without any 0
        const int r_hi = y1 + R;
        for (int r = y0 - R; r < 0; r++) {
            uint32_t C = FULL, L = FULL;
            if (phantom && r == -1) {
                if (k == 0) C = 0xfffffffeu;
                if (k == 1) L = 0xfffffffeu;
            }
            feed(C, L, FULL);
        }
        const int ra = max(y0 - R, 0), rb = min(r_hi, H);    // rows that exist
        const int rm = min(max(y0 + R, ra), rb);             // first row whose feed completes an output row
        const uint32_t *row = src + ra * wpr + k;
        for (int r = ra; r < rm; r++, row += wpr) feed((row[0] ^ inv) | padC, (row[oL] ^ inv) | mL, (row[oR] ^ inv) | mR);
        uint32_t *out = dst + (rm - R) * wpr + k;
        for (int r = rm; r < rb; r++, row += wpr, out += wpr) {
            feed((row[0] ^ inv) | padC, (row[oL] ^ inv) | mL, (row[oR] ^ inv) | mR);
            emit(out, r - R);
        }
        for (int r = rb; r < r_hi; r++) { // below the image (last strip only)
            feed(FULL, FULL, FULL);
            if (r - R >= y0) emit(dst + (r - R) * wpr + k, r - R);
        }
    }
    hz_next |= hzacc != 0u;
}

// "pixel > t" for 32 pixels held in nine aligned 32-bit words (row misaligned by `a` bytes).  Byte-wise compare in
// three integer ops per four pixels (LOW, t < 128: msb(px) | msb(low7(px) + 127 - t); else msb(px) & msb(low7(px) +
// 255 - t)); the msbs of two groups are interleaved (bits 0, 4, 8, ...) and gathered into one byte of the plane
// by ONE multiply: bit b lands at b + {24, 17, 10, 3}, and no two partial products meet in the top byte.
template <bool LOW>
__device__ __forceinline__ uint32_t threshold32(const uint32_t (&raw)[9], uint32_t a, uint32_t addc)
{
    uint32_t byte4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t m[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t px = __funnelshift_r(raw[2 * j + h], raw[2 * j + h + 1], 8 * a);
            const uint32_t s7 = (px & 0x7f7f7f7fu) + addc;  // no carry across bytes
            m[h] = (LOW ? (s7 | px) : (s7 & px)) & 0x80808080u;
        }
        byte4[j] = ((m[0] >> 7) | (m[1] >> 3)) * 0x01020408u;  // top byte = pixels 8j .. 8j+7
    }
    return __byte_perm(__byte_perm(byte4[0], byte4[1], 0x0073), __byte_perm(byte4[2], byte4[3], 0x7300), 0x7610);
}
