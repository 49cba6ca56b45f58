// Shared device helpers for the maze_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "maze_b200.h"

typedef long long i64;
typedef unsigned long long u64;

#define MAZE_CTA 256 /* threads per CTA == MAZE_TILE_WORDS: one thread per word or 8 warps x 32 words */
#define FULL 0xffffffffu

extern thread_local char maze_err_buf[256];
void maze_set_err(cudaError_t e, const char *where);

#define MAZE_LAUNCH_CHECK(where)                      \
    do {                                              \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) {                     \
            maze_set_err(e__, where);                 \
            return MAZE_ERR_CUDA;                     \
        }                                             \
    } while (0)

// ---- launch accounting / optional per-kernel CUDA-event timing (maze_prof.cu) ---------------------
enum MazeKernelId {
    KID_THRESHOLD_PACK = 0, KID_COMPARE_PACK, KID_UNPACK_MASK, KID_MORPH_PASS, KID_PLANE_HAS_ZERO, KID_EDT_COLS,
    KID_EDT_ROWS, KID_CCL_INIT, KID_CCL_UNION, KID_CCL_FLATTEN, KID_TILE_SCAN, KID_CCL_ASSIGN, KID_CCL_WRITE,
    KID_BORDER_MARK, KID_LABEL_ZERO, KID_LABEL_COUNT, KID_MAX_LABEL, KID_PROPS_INIT, KID_PROPS_ACCUMULATE,
    KID_PROPS_FINISH, KID_PROPS_HIGH_ORDER, KID_MERGE_LABELS, KID_SYNTH, KID_PROPS_RUNS, KID_PROPS_RUNS_HIGH, KID_VIGNETTE_FUSED, KID_COUNT_SCAN, KID_PROPS_FINISH_STAGED, KID_SCATTER_COUNTS, KID_LABEL_SHAPE, KID_BAND_FRONT, KID_BAND_LABEL, KID_BAND_LABEL_BIG, KID_BAND_WRITE, KID_BAND_ZERO, KID_WIDE_VDIST, KID_WIDE_ROWS, KID_GL_PREFIX, KID_GL_LINK, KID_GL_RANK, KID_GL_APPLY, KID_MERGE_WINDOWED, KID_MERGE_PREPARE, KID_COUNT
};
void maze_prof_begin(int kid, cudaStream_t s);
void maze_prof_end(int kid, cudaStream_t s);

// launch a kernel: counted, optionally timed, checked
#define MAZE_KERNEL(kid, s, ...)                      \
    do {                                              \
        maze_prof_begin(kid, s);                      \
        __VA_ARGS__;                                  \
        maze_prof_end(kid, s);                        \
        MAZE_LAUNCH_CHECK(#kid);                      \
    } while (0)

#define MAZE_CUDA(call, where)                        \
    do {                                              \
        cudaError_t e__ = (call);                     \
        if (e__ != cudaSuccess) {                     \
            maze_set_err(e__, where);                 \
            return MAZE_ERR_CUDA;                     \
        }                                             \
    } while (0)

// Chord table of one morphology pass from its threshold code (maze_morph.cu): t >= 0 is the squared-distance
// threshold of the EDT compare (closed disk of d2 <= t, scipy's phantom pixel applies); t = MAZE_FOOTPRINT_T(id)
// selects a footprint registered with maze_footprint_register (plain binary erosion / dilation, no phantom).
// Returns false for an invalid code.
bool maze_pass_table(int t, int *R, int *w /* MAZE_MAX_DISK_RADIUS + 1 */, int *use_phantom);

// bits of word k of a row of width w that are real pixels
__device__ __forceinline__ uint32_t valid_mask(int w, int k)
{
    int rem = w - 32 * k;
    return rem >= 32 ? FULL : ((1u << rem) - 1u);
}

// first bit of the horizontal run (inside this word) that contains bit b of m
__device__ __forceinline__ int run_start_in_word(uint32_t m, int b)
{
    uint32_t z = ~m & ((1u << b) - 1u);
    return z ? 32 - __clz(z) : 0;
}

__device__ __forceinline__ int ld_volatile(const int *p) { return *(const volatile int *)p; }

// Union-find on per-vignette pixel indices; the root of a set is its smallest index, i.e. the
// component's first pixel in raster order.
__device__ __forceinline__ int uf_find(const int *P, int n)
{
    int p = ld_volatile(P + n);
    while (p != n) {
        n = p;
        p = ld_volatile(P + n);
    }
    return n;
}

__device__ __forceinline__ void uf_union(int *P, int a, int b)
{
    while (true) {
        a = uf_find(P, a);
        b = uf_find(P, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(P + a, b);
        if (old == a) return;
        a = old;
    }
}

// exclusive scan of one int per thread over a 256-thread CTA; returns the exclusive prefix,
// *total receives the CTA sum.  smem: >= 9 ints.
__device__ __forceinline__ int cta_exclusive_scan(int v, int *smem, int *total)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int nw = blockDim.x >> 5;
        int s = lane < nw ? smem[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULL, si, d);
            if (lane >= d) si += t;
        }
        if (lane < nw) smem[lane] = si - s;
        if (lane == nw - 1) smem[nw] = si;
    }
    __syncthreads();
    int res = smem[warp] + inc - v;
    *total = smem[blockDim.x >> 5];
    __syncthreads();
    return res;
}

// bits of "pixel > t" for the 32 pixels of word k of row y (row-major uint8 image of width W): nine aligned
// 32-bit loads, byte alignment by funnel shift, byte-wise compare as three integer ops on four pixels
// (t < 128: px > t <=> msb(px) | msb(low7(px) + 127 - t); t >= 128: msb(px) & msb(low7(px) + 255 - t)), the
// four msbs gathered into a nibble by a multiply.  Never reads past the aligned word holding the row's last pixel.
__device__ __forceinline__ uint32_t threshold_word32(const uint8_t *base, int y, int k, int W, int t)
{
    if (t < 0) return valid_mask(W, k);
    if (t >= 255) return 0u;
    const uint32_t addc = (uint32_t)((t < 128 ? 127 - t : 255 - t) & 0x7f) * 0x01010101u;
    const bool hi_t = t >= 128;
    const int nvalid = min(32, W - 32 * k);
    const uint8_t *p = base + (size_t)y * W + 32 * k;
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *q = (const uint32_t *)(p - a);
    uint32_t word = 0, lo = __ldg(q);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (4 * i < nvalid) {
            uint32_t hi = (4 * (i + 1) < (int)a + nvalid) ? __ldg(q + i + 1) : 0u;  // never past the last pixel's word
            uint32_t px = __funnelshift_r(lo, hi, 8 * a);
            uint32_t s7 = (px & 0x7f7f7f7fu) + addc;
            uint32_t m = (hi_t ? (s7 & px) : (s7 | px)) & 0x80808080u;
            word += (((m >> 7) * 0x01020408u) >> 24) << (4 * i);
            lo = hi;
        }
    }
    return word & valid_mask(W, k);
}

struct TileCtx {
    maze_vignette_t v;
    int img;
    int word0;
    int nwords; // words in the vignette
};

__device__ __forceinline__ TileCtx load_tile(const maze_vignette_t *vig, const maze_tile_t *tiles)
{
    TileCtx c;
    maze_tile_t t = tiles[blockIdx.x];
    c.img = t.img;
    c.word0 = t.word0;
    c.v = vig[t.img];
    c.nwords = c.v.h * c.v.wpr;
    return c;
}
