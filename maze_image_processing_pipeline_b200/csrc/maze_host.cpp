// Host-side helpers of the C-ABI (no CUDA): multi-threaded packing of many small vignettes into the flat
// pinned staging buffer that is then uploaded with ONE copy (and the reverse scatter is not needed:
// results are numpy views of the flat download).
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "maze_b200.h"

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
            for (int i = lo; i < hi; i++) memcpy((char *)dst + dst_off[i], srcs[i], (size_t)nbytes[i]);
        });
    }
    for (auto &x : th) x.join();
    return MAZE_OK;
}
