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

// Asynchronous form of maze_host_pack: the copy threads start at once and the call returns a job handle; the caller
// goes on with its bookkeeping (geometry, descriptors, launches of the previous batch) and waits for the job right
// before it uploads the staging buffer.  The descriptor arrays are copied into the job; the SOURCE arrays and the
// destination must stay alive until maze_host_pack_wait returns.
struct PackJob {
    std::vector<const void *> srcs;
    std::vector<int64_t> nbytes, dst_off;
    std::vector<std::thread> threads;
};

extern "C" void *maze_host_pack_start(const void *const *srcs, const int64_t *nbytes, const int64_t *dst_off, int n,
                                      void *dst, int n_threads)
{
    if (n < 0 || (n > 0 && (!srcs || !nbytes || !dst_off || !dst))) return nullptr;
    PackJob *job = new PackJob;
    if (n == 0) return job;
    job->srcs.assign(srcs, srcs + n);
    job->nbytes.assign(nbytes, nbytes + n);
    job->dst_off.assign(dst_off, dst_off + n);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += nbytes[i];
    std::vector<int> cut(n_threads + 1, n);
    cut[0] = 0;
    int64_t acc = 0;
    int t = 1;
    for (int i = 0; i < n && t < n_threads; i++) {
        acc += nbytes[i];
        if (acc >= total * t / n_threads) cut[t++] = i + 1;
    }
    for (int k = 0; k < n_threads; k++) {
        const int lo = cut[k], hi = cut[k + 1];
        if (lo >= hi) continue;
        job->threads.emplace_back([job, dst, lo, hi]() {
            for (int i = lo; i < hi; i++)
                stream_copy((char *)dst + job->dst_off[i], (const char *)job->srcs[i], (size_t)job->nbytes[i]);
            fence_stores();
        });
    }
    return job;
}

extern "C" int maze_host_pack_wait(void *handle)
{
    if (!handle) return MAZE_ERR_BADARG;
    PackJob *job = (PackJob *)handle;
    for (auto &t : job->threads) t.join();
    delete job;
    return MAZE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Host side of the compact result transport: the label image of a vignette crosses PCIe as its RUN LIST
// (maze_run_t {y, x0, x1, label}, 8 bytes per run, written by maze_band_stage) and is expanded here on demand
// into what the reference's stage returns (bool mask + int32 label image, loki/pipeline.py:459) or into the
// padded crop of one object (what FindRegions / ExtractROI consume, loki/pipeline.py:589-602).
// ---------------------------------------------------------------------------------------------------------
// memset + painting in place: any order of the runs
static void expand_one_unordered(const maze_run_t *runs, const maze_band_out_t *band_out, int b_lo, int b_hi, int h, int w,
                                 uint8_t *mask, int32_t *labels)
{
    const size_t npx = (size_t)h * (size_t)w;
    if (mask) memset(mask, 0, npx);
    if (labels) memset(labels, 0, npx * sizeof(int32_t));
    for (int b = b_lo; b < b_hi; b++) {
        const maze_band_out_t o = band_out[b];
        if (o.base < 0) continue;
        const maze_run_t *r = runs + o.base;
        for (int i = 0; i < o.n_runs; i++) {
            if ((int)r[i].y >= h || r[i].x1 < r[i].x0 || (int)r[i].x1 >= w) continue; // (never written outside the vignette)
            const size_t off = (size_t)r[i].y * (size_t)w + r[i].x0;
            const int len = (int)r[i].x1 - (int)r[i].x0 + 1;
            if (mask) memset(mask + off, 1, (size_t)len);
            if (labels) {
                const int32_t lab = r[i].label;
                int32_t *d = labels + off;
                for (int x = 0; x < len; x++) d[x] = lab;
            }
        }
    }
}

static void expand_one(const maze_run_t *runs, const maze_band_out_t *band_out, int b_lo, int b_hi, int h, int w,
                       uint8_t *mask, int32_t *labels)
{
    {
        // (the streaming composition below relies on raster order; the run list is ~1 % of the output bytes: check)
        size_t last = 0;
        for (int b = b_lo; b < b_hi; b++) {
            const maze_band_out_t o = band_out[b];
            if (o.base < 0) continue;
            const maze_run_t *r = runs + o.base;
            for (int i = 0; i < o.n_runs; i++) {
                const size_t p = (size_t)r[i].y * (size_t)w + r[i].x0;
                if (p < last || r[i].x1 < r[i].x0 || (int)r[i].x1 >= w || (int)r[i].y >= h) {
                    expand_one_unordered(runs, band_out, b_lo, b_hi, h, w, mask, labels);
                    return;
                }
                last = p;
            }
        }
    }
    // The runs of a vignette are in raster order (bands in order, raster order inside a band), and a vignette is one
    // contiguous block of h * w pixels: the output is composed chunk by chunk in a buffer that stays in L1 / L2 (zeros,
    // then the pieces of the runs that fall into the chunk) and leaves with non-temporal stores -- the arrays are
    // written ONCE, no line is read for ownership (memset + painting in place costs the memory system twice the bytes).
    const size_t npx = (size_t)h * (size_t)w;
    constexpr size_t CH = 8192; // pixels per chunk: 8 KB of mask + 32 KB of labels
    alignas(64) uint8_t mb[CH];
    alignas(64) int32_t lb[CH];
    int b = b_lo, i = 0;
    size_t p0 = 0, p1 = 0; // current run as flat pixel range [p0, p1)
    int32_t lab = 0;
    bool have = false;
    auto next_run = [&]() {
        have = false;
        while (b < b_hi) {
            const maze_band_out_t o = band_out[b];
            if (o.base < 0 || i >= o.n_runs) { b++; i = 0; continue; }
            const maze_run_t r = runs[o.base + i++];
            p0 = (size_t)r.y * (size_t)w + r.x0;
            p1 = (size_t)r.y * (size_t)w + r.x1 + 1;
            lab = r.label;
            have = true;
            return;
        }
    };
    next_run();
    for (size_t c0 = 0; c0 < npx; c0 += CH) {
        const size_t c1 = c0 + CH < npx ? c0 + CH : npx, n = c1 - c0;
        if (mask) memset(mb, 0, n);
        if (labels) memset(lb, 0, n * sizeof(int32_t));
        while (have && p0 < c1) {
            const size_t a = (p0 > c0 ? p0 : c0) - c0, e = (p1 < c1 ? p1 : c1) - c0;
            if (e > a) { // (a run that lies before the chunk -- runs out of raster order -- is dropped, not written wildly)
                if (mask) memset(mb + a, 1, e - a);
                if (labels)
                    for (size_t x = a; x < e; x++) lb[x] = lab;
            }
            if (p1 <= c1) next_run(); else break;
        }
        if (mask) stream_copy((char *)(mask + c0), (const char *)mb, n);
        if (labels) stream_copy((char *)(labels + c0), (const char *)lb, n * sizeof(int32_t));
    }
    fence_stores();
}

extern "C" int maze_host_expand(const maze_run_t *runs, const maze_band_out_t *band_out, const int32_t *band_lo,
                                const int32_t *band_hi, const int32_t *h, const int32_t *w, int n,
                                uint8_t *const *mask_dst, int32_t *const *label_dst, int n_threads)
{
    if (n <= 0) return MAZE_OK;
    if (!runs || !band_out || !band_lo || !band_hi || !h || !w) return MAZE_ERR_BADARG;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    int64_t total = 0;
    for (int i = 0; i < n; i++) total += (int64_t)h[i] * w[i];
    auto work = [=](int lo, int hi) {
        for (int i = lo; i < hi; i++)
            expand_one(runs, band_out, band_lo[i], band_hi[i], h[i], w[i], mask_dst ? mask_dst[i] : nullptr,
                       label_dst ? label_dst[i] : nullptr);
    };
    if (n_threads == 1 || n == 1 || total < (1 << 20)) {
        work(0, n);
        return MAZE_OK;
    }
    std::vector<int> cut(n_threads + 1, n);
    cut[0] = 0;
    int64_t acc = 0;
    int t = 1;
    for (int i = 0; i < n && t < n_threads; i++) {
        acc += (int64_t)h[i] * w[i];
        if (acc >= total * t / n_threads) cut[t++] = i + 1;
    }
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; k++)
        if (cut[k] < cut[k + 1]) th.emplace_back(work, cut[k], cut[k + 1]);
    for (auto &x : th) x.join();
    return MAZE_OK;
}

// rows [r0, r1) x columns [c0, c1) of one vignette (inside the image) as a contiguous crop.  only_label > 0: the
// mask holds that object alone (RegionProperties.image), otherwise every foreground pixel.
extern "C" int maze_host_expand_crop(const maze_run_t *runs, const maze_band_out_t *band_out, int band_lo, int band_hi,
                                     int rpb, int r0, int r1, int c0, int c1, int only_label, uint8_t *mask_dst,
                                     int32_t *label_dst)
{
    if (!runs || !band_out || rpb < 1 || r0 < 0 || c0 < 0 || r1 < r0 || c1 < c0) return MAZE_ERR_BADARG;
    const int cw = c1 - c0;
    const size_t npx = (size_t)(r1 - r0) * (size_t)cw;
    if (mask_dst) memset(mask_dst, 0, npx);
    if (label_dst) memset(label_dst, 0, npx * sizeof(int32_t));
    if (npx == 0) return MAZE_OK;
    int b_first = band_lo + r0 / rpb, b_last = band_lo + (r1 - 1) / rpb;
    if (b_last >= band_hi) b_last = band_hi - 1;
    for (int b = b_first; b <= b_last; b++) {
        const maze_band_out_t o = band_out[b];
        if (o.base < 0) return MAZE_ERR_CAPACITY; // this vignette has no run list
        const maze_run_t *r = runs + o.base;
        int lo = 0, hi = o.n_runs; // first run with y >= r0 (runs are in raster order)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)r[mid].y < r0) lo = mid + 1; else hi = mid;
        }
        for (int i = lo; i < o.n_runs && (int)r[i].y < r1; i++) {
            if (only_label > 0 && (int)r[i].label != only_label) continue;
            const int a = (int)r[i].x0 > c0 ? (int)r[i].x0 : c0, e = (int)r[i].x1 + 1 < c1 ? (int)r[i].x1 + 1 : c1;
            if (a >= e) continue;
            const size_t off = (size_t)((int)r[i].y - r0) * (size_t)cw + (size_t)(a - c0);
            if (mask_dst) memset(mask_dst + off, 1, (size_t)(e - a));
            if (label_dst) {
                const int32_t lab = r[i].label;
                for (int x = 0; x < e - a; x++) label_dst[off + x] = lab;
            }
        }
    }
    return MAZE_OK;
}
