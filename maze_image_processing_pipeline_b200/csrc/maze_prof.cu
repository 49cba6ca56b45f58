// Launch accounting and optional per-kernel CUDA-event timing for the roofline report.
// Every kernel of the library is launched through MAZE_KERNEL (maze_common.cuh): the launch is always
// counted; when timing is enabled a start/stop event pair is recorded around it on the launching
// stream and maze_prof_collect() sums the elapsed times per kernel.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "maze_common.cuh"

static const char *const k_names[KID_COUNT] = {
    "k_threshold_pack", "k_compare_pack", "k_unpack_mask", "k_morph_pass", "k_plane_has_zero", "k_edt_cols",
    "k_edt_rows", "k_ccl_init", "k_ccl_union", "k_ccl_flatten", "k_tile_scan", "k_ccl_assign", "k_ccl_write",
    "k_border_mark", "k_label_zero", "k_label_count", "k_max_label", "k_props_init", "k_props_accumulate",
    "k_props_finish", "k_props_high_order", "k_merge_labels", "k_synth", "k_props_runs", "k_props_runs_high", "k_vignette_fused", "k_count_scan", "k_props_finish_staged", "k_scatter_counts", "k_label_shape", "k_band_front", "k_band_label", "k_band_label_big", "k_band_write", "k_band_zero", "k_wide_vdist", "k_wide_rows", "k_gl_prefix", "k_gl_link", "k_gl_rank", "k_gl_apply", "k_merge_windowed", "k_mw_prepare"};

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_enabled{0};
static std::mutex g_mu;
struct ProfRec { int kid; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_free;
static thread_local cudaEvent_t t_pending_start = nullptr;

static cudaEvent_t get_event()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_free.empty()) {
        cudaEvent_t e = g_free.back();
        g_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void maze_prof_begin(int kid, cudaStream_t s)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_enabled.load(std::memory_order_relaxed)) return;
    t_pending_start = get_event();
    cudaEventRecord(t_pending_start, s);
}

void maze_prof_end(int kid, cudaStream_t s)
{
    if (!t_pending_start) return;
    cudaEvent_t b = get_event();
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back({kid, t_pending_start, b});
    t_pending_start = nullptr;
}

extern "C" long long maze_launch_count(void) { return g_launches.load(); }
void maze_prof_add_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int maze_prof_is_enabled(void) { return g_enabled.load(std::memory_order_relaxed) ? 1 : 0; }
extern "C" int maze_prof_kernel_count(void) { return KID_COUNT; }
extern "C" const char *maze_prof_kernel_name(int kid) { return (kid >= 0 && kid < KID_COUNT) ? k_names[kid] : ""; }
extern "C" int maze_prof_enable(int on)
{
    g_enabled.store(on ? 1 : 0);
    return MAZE_OK;
}

// Waits for all recorded events, adds the elapsed milliseconds / launch counts per kernel id into
// ms[0..n) / counts[0..n) (host arrays the caller zeroed) and forgets the records.  Launches of one
// kernel id that overlap in time (the concurrent size classes of the fused stage) are counted once:
// ms is the length of the UNION of the [start, stop] intervals of that id.
extern "C" int maze_prof_collect(double *ms, long long *counts, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_recs.empty()) return MAZE_OK;
    std::vector<std::vector<std::pair<double, double>>> iv((size_t)n);
    cudaEvent_t ref = g_recs[0].a;
    for (auto &r : g_recs) {
        MAZE_CUDA(cudaEventSynchronize(r.b), "prof sync");
        float t0 = 0.f, t1 = 0.f;
        MAZE_CUDA(cudaEventElapsedTime(&t0, ref, r.a), "prof elapsed");
        MAZE_CUDA(cudaEventElapsedTime(&t1, ref, r.b), "prof elapsed");
        if (r.kid >= 0 && r.kid < n) {
            iv[(size_t)r.kid].push_back({(double)t0, (double)t1});
            counts[r.kid] += 1;
        }
    }
    for (int k = 0; k < n; k++) {
        auto &v = iv[(size_t)k];
        std::sort(v.begin(), v.end());
        double lo = 0, hi = -1;
        for (auto &p : v) {
            if (hi < lo || p.first > hi) {
                if (hi >= lo) ms[k] += hi - lo;
                lo = p.first;
                hi = p.second;
            } else if (p.second > hi) {
                hi = p.second;
            }
        }
        if (hi >= lo) ms[k] += hi - lo;
    }
    for (auto &r : g_recs) {
        g_free.push_back(r.a);
        g_free.push_back(r.b);
    }
    g_recs.clear();
    return MAZE_OK;
}
