// One asynchronous step of the fused LOKI stage as a single C call: zeroing of the counters, the
// vignette-resident kernel on the lane stream, the per-operator chain for the vignettes it cannot hold on
// the side stream, label offsets, feature rows, and the read-back of the counters into pinned memory.
// Same sequence as stage.LokiSegmentationStage used to issue from Python (a dozen binding calls and torch
// ops per step); keeping it in C makes the host cost of a step a few tens of microseconds.
#include "maze_common.cuh"

__global__ void k_scatter_counts(const int32_t *__restrict__ sub_lab_off, const int32_t *__restrict__ idx, int n,
                                 int32_t *__restrict__ n_labels)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) n_labels[idx[i]] = sub_lab_off[i + 1] - sub_lab_off[i];
}

struct StepEvents {
    int device;
    cudaEvent_t ev[3];
};

static StepEvents *step_events()
{
    static thread_local StepEvents pool[16];
    static thread_local int n_pool = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < n_pool; i++)
        if (pool[i].device == dev) return &pool[i];
    if (n_pool >= 16) return nullptr;
    StepEvents *e = &pool[n_pool];
    e->device = dev;
    for (int k = 0; k < 3; k++)
        if (cudaEventCreateWithFlags(&e->ev[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    n_pool++;
    return e;
}

extern "C" int maze_stage_step(const maze_step_args_t *a, void *lane_stream, void *side_stream)
{
    cudaStream_t lane = (cudaStream_t)lane_stream, side = (cudaStream_t)side_stream;
    const int n = a->n_img;
    if (n <= 0) return MAZE_OK;
    StepEvents *se = step_events();
    if (!se) return MAZE_ERR_CUDA;
    int32_t *n_labels = a->counts, *fallback = a->counts + n, *acc_base = a->counts + 2 * n;
    MAZE_CUDA(cudaMemsetAsync(a->counts, 0, sizeof(int32_t) * 2 * (size_t)n, lane), "step counters");
    MAZE_CUDA(cudaMemsetAsync(acc_base, 0xff, sizeof(int32_t) * (size_t)n, lane), "step acc_base");
    int rc;
    const bool left = a->left_n > 0;
    if (left) {
        MAZE_CUDA(cudaEventRecord(se->ev[0], lane), "step fork");
        MAZE_CUDA(cudaStreamWaitEvent(side, se->ev[0], 0), "step fork wait");
        // the final plane of the chain must be `bits`: it is plane a after an even number of passes, else b
        uint32_t *pa = (a->n_pass & 1) ? a->scratch_plane : a->bits;
        uint32_t *pb = (a->n_pass & 1) ? a->bits : a->scratch_plane;
        uint32_t *fin = nullptr;
        rc = maze_front_chain(a->image, a->left_vig, a->left_n, a->left_tiles, a->left_n_tiles, a->t_int, a->n_pass,
                              a->pass_t, a->pass_invert, pa, pb, a->scratch_flags, a->scratch_flags + a->left_n,
                              a->scratch_parent, a->labels, a->scratch_tile_scan, a->scratch_lab_off, a->mask, &fin,
                              side);
        if (rc != MAZE_OK) return rc;
        if (fin != a->bits) return MAZE_ERR_BADARG;
        MAZE_KERNEL(KID_SCATTER_COUNTS, side,
                    k_scatter_counts<<<(a->left_n + 127) / 128, 128, 0, side>>>(a->scratch_lab_off, a->left_idx,
                                                                                a->left_n, n_labels));
        MAZE_CUDA(cudaEventRecord(se->ev[1], side), "step join");
    }
    if (a->bands) {
        const bool dense = !(a->step_flags & MAZE_STEP_COMPACT);
        rc = maze_band_stage(a->image, a->intensity, a->vig, n, a->bands, a->n_bands, a->band_off, a->t_int, a->n_pass,
                             a->pass_t, a->pass_invert, a->halo, a->flags, a->bits, a->runs, a->run_pix, a->run_stats,
                             a->run_cap, a->band_out, dense ? a->mask : nullptr, dense ? a->labels : nullptr, n_labels,
                             fallback, acc_base, a->band_counters, a->big_list, a->stage_cap, a->acc_stage, a->hi_stage,
                             a->ext_stage, a->total_px, a->huge_host, a->n_huge, a->huge_px, a->gl_scratch, a->band_done, a->clear_border, a->min_area, lane);
    } else {
        rc = maze_vignette_stage(a->image, a->intensity, a->vig, a->img_list, a->class_off, a->t_int, a->n_pass, a->pass_t,
                                 a->pass_invert, a->flags, a->bits, a->mask, a->labels, n_labels, fallback, acc_base,
                                 a->stage_counter, a->stage_cap, a->acc_stage, a->hi_stage, a->ext_stage, lane);
    }
    if (rc != MAZE_OK) return rc;
    if (left) MAZE_CUDA(cudaStreamWaitEvent(lane, se->ev[1], 0), "step join wait");
    rc = maze_count_scan(n_labels, n, a->lab_off, lane);
    if (rc != MAZE_OK) return rc;
    rc = maze_props_finish_staged(a->acc_stage, a->hi_stage, a->ext_stage, acc_base, a->lab_off, n, a->stage_cap,
                                  a->intensity ? 1 : 0, a->flags, a->table, lane);
    if (rc != MAZE_OK) return rc;
    if (left) {
        MAZE_CUDA(cudaEventRecord(se->ev[2], lane), "step fork 2");
        MAZE_CUDA(cudaStreamWaitEvent(side, se->ev[2], 0), "step fork 2 wait");
        rc = maze_regionprops(a->labels, a->bits, a->intensity, a->vig, n, a->left_tiles_full, a->left_n_tiles_full,
                              a->lab_off, a->stage_cap, a->scratch_acc, a->scratch_ext, a->table,
                              (a->flags & MAZE_RP_HIGH_ORDER) | MAZE_RP_RUNS, acc_base, side);
        if (rc != MAZE_OK) return rc;
    }
    if (a->counts_host) {
        MAZE_CUDA(cudaMemcpyAsync(a->counts_host, a->counts, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyDeviceToHost,
                                  lane), "step readback");
        MAZE_CUDA(cudaMemcpyAsync(a->counts_host + 3 * (size_t)n, a->lab_off + n, sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, lane), "step readback total");
        if (a->bands)
            MAZE_CUDA(cudaMemcpyAsync(a->counts_host + 3 * (size_t)n + 1, a->band_counters + 1, sizeof(int32_t),
                                      cudaMemcpyDeviceToHost, lane), "step readback runs");
    }
    return MAZE_OK;
}

void maze_prof_add_launches(long long n); // maze_prof.cu
int maze_prof_is_enabled(void);

// The same step as a CUDA GRAPH: a batch that is one frame (BASELINE.json configs[3]) is some twenty small launches, and
// the host time of issuing them (~0.1 ms) exceeds the GPU time of the frame; replaying the captured step costs one launch.
// *exec == NULL: the step is captured on the lane stream (relaxed mode), instantiated, launched, and the handle and the
// number of kernel launches it holds are returned; otherwise the graph is launched.  A graph replays the ARGUMENTS it was
// captured with: the caller keys its handles by the argument block (pointers, sizes, flags -- and the band plan behind
// huge_host, which is read on the host at capture time).  Steps with leftover vignettes (side-stream work that joins
// later) and steps under the per-kernel profiler are not captured: *n_launches = -1 tells the caller to stop asking.
extern "C" int maze_stage_step_graph(const maze_step_args_t *a, void *lane_stream, void *side_stream, void **exec,
                                     int *n_launches)
{
    cudaStream_t lane = (cudaStream_t)lane_stream;
    if (!exec || !n_launches) return MAZE_ERR_BADARG;
    if (*exec) {
        MAZE_CUDA(cudaGraphLaunch((cudaGraphExec_t)*exec, lane), "step graph launch");
        maze_prof_add_launches(*n_launches);
        return MAZE_OK;
    }
    if (a->left_n > 0 || maze_prof_is_enabled()) {
        *n_launches = -1;
        return maze_stage_step(a, lane_stream, side_stream);
    }
    const long long before = maze_launch_count();
    if (cudaStreamBeginCapture(lane, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        *n_launches = -1;
        return maze_stage_step(a, lane_stream, side_stream);
    }
    const int rc = maze_stage_step(a, lane_stream, side_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(lane, &graph);
    const long long captured = maze_launch_count() - before;
    maze_prof_add_launches(-captured); // (nothing ran yet)
    cudaGraphExec_t ge = nullptr;
    if (rc != MAZE_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&ge, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        *n_launches = -1;
        if (rc != MAZE_OK) return rc;
        return maze_stage_step(a, lane_stream, side_stream);
    }
    cudaGraphDestroy(graph);
    *exec = (void *)ge;
    *n_launches = (int)captured;
    MAZE_CUDA(cudaGraphLaunch(ge, lane), "step graph launch");
    maze_prof_add_launches(captured);
    return MAZE_OK;
}

extern "C" int maze_graph_destroy(void *exec)
{
    if (exec) cudaGraphExecDestroy((cudaGraphExec_t)exec);
    return MAZE_OK;
}
