// Host-side helpers of the C-ABI (no CUDA): multi-threaded packing of many small vignettes into the flat
// pinned staging buffer that is then uploaded with ONE copy (and the reverse scatter is not needed:
// results are numpy views of the flat download).
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "maze_b200.h"

// Copy into the pinned staging buffer with non-temporal stores: the destination is read next by the DMA engine,
// never by this core, so a plain memcpy would first fetch every destination line (read-for-ownership) and push
// the source out of the cache -- a third of the memory traffic of the pack, on a host whose memory system is
// also absorbing the 50 GB/s download of the previous batch.
static void stream_copy(char *dst, const char *src, size_t n)
{
#if defined(__SSE2__)
    while (n && ((uintptr_t)dst & 15)) { *dst++ = *src++; n--; }
    size_t blocks = n / 64;
    for (size_t b = 0; b < blocks; b++) {
        __m128i a0 = _mm_loadu_si128((const __m128i *)(src + 0)), a1 = _mm_loadu_si128((const __m128i *)(src + 16));
        __m128i a2 = _mm_loadu_si128((const __m128i *)(src + 32)), a3 = _mm_loadu_si128((const __m128i *)(src + 48));
        _mm_stream_si128((__m128i *)(dst + 0), a0);
        _mm_stream_si128((__m128i *)(dst + 16), a1);
        _mm_stream_si128((__m128i *)(dst + 32), a2);
        _mm_stream_si128((__m128i *)(dst + 48), a3);
        src += 64; dst += 64;
    }
    n -= blocks * 64;
    if (n) memcpy(dst, src, n);
#else
    memcpy(dst, src, n);
#endif
}

static void fence_stores()
{
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

extern "C" int maze_host_pack(const void *const *srcs, const int64_t *nbytes, const int64_t *dst_off, int n,
                              void *dst, int n_threads)
{
    if (n <= 0) return MAZE_OK;
    if (!srcs || !nbytes || !dst_off || !dst) return MAZE_ERR_BADARG;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += nbytes[i];
    if (total < (1 << 20) || n_threads == 1) {
        for (int i = 0; i < n; i++) memcpy((char *)dst + dst_off[i], srcs[i], (size_t)nbytes[i]);
        return MAZE_OK;
    }
    // contiguous ranges of vignettes with about the same number of bytes per thread
    std::vector<int> cut(n_threads + 1, n);
    cut[0] = 0;
    int64_t acc = 0;
    int t = 1;
    for (int i = 0; i < n && t < n_threads; i++) {
        acc += nbytes[i];
        if (acc >= total * t / n_threads) cut[t++] = i + 1;
    }
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; k++) {
        int lo = cut[k], hi = cut[k + 1];
        if (lo >= hi) continue;
        th.emplace_back([=]() {
            for (int i = lo; i < hi; i++) stream_copy((char *)dst + dst_off[i], (const char *)srcs[i], (size_t)nbytes[i]);
            fence_stores();
        });
    }
    for (auto &x : th) x.join();
    return MAZE_OK;
}
