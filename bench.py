#!/usr/bin/env python
"""Benchmark of the LOKI re-segmentation stage (BASELINE.json metric: vignettes/s and MPix/s for
seg + CCL + regionprops, HBM % of peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): 100 000 synthetic variable-size LOKI vignettes (H, W
independently log-uniform in [64, 1024]), processed in batches of --batch vignettes; one STEP is one
pass of the whole chain (threshold 40 -> isotropic opening r=1 -> isotropic closing r=2 -> 8-connected
labelling -> regionprops, mask bytes written) over one batch.  Batch b holds vignettes
[b*B, (b+1)*B) of the job; rank r of N works on batches r, r+N, ... (weak scaling, no collective on
the data path).  Every batch is > L2 (about 490 MB of pixels at the default 4096 vignettes), so no L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

JOB_VIGNETTES = 100_000
MAX_RESIDENT = 24    # distinct resident batches per rank (the job has 25 batches of 4096)
SIZE_SEED = 1        # C2 uses seed 1
PIXEL_SEED = 20261018
THRESHOLD, R_OPEN, R_CLOSE = 40, 1, 2

# algorithmic bytes per pixel of each kernel (inputs it must read + outputs it must write)
KERNEL_BYTES_PER_PX = {
    "k_threshold_pack": 1.125, "k_morph_pass": 0.25, "k_unpack_mask": 1.125, "k_ccl_init": 0.125,
    "k_ccl_union": 0.125, "k_ccl_flatten": 0.125, "k_ccl_assign": 0.125, "k_ccl_write": 4.125,
    "k_props_accumulate": 5.0, "k_props_high_order": 4.0, "k_label_zero": 8.0, "k_label_count": 4.0,
    "k_vignette_fused": 6.125, "k_props_runs": 5.0, "k_props_runs_high": 4.0,
    # band front: 1 R image + 1 W mask + 4 W label image (the zero fill; the labelling kernel rewrites only the
    # foreground runs) + 1/8 W bit plane
    "k_band_front": 6.125,
}
STAGE_BYTES_PER_PX = 6.0  # 1 R image + 1 W mask + 4 W labels


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per pixel of `kernel` from the committed ncu capture
    (profiles/ncu_traffic.json, written from an ncu launch list of THIS bench command), or None when no capture of
    that kernel is on file: the bench never invents the number."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p)).get(kernel)
        return None if rec is None else (float(rec["dram_bytes_per_px"]), rec.get("source", p))
    except Exception:
        return None


def release(*objs):
    """Drop device workspaces of stage objects that are no longer needed (each holds several GB per lane)."""
    import gc
    import torch
    del objs
    gc.collect()
    torch.cuda.empty_cache()


def windowed_steps(stage, items):
    """Run the stage over resident (batch, image) pairs with the read-back of a step checked n_lanes - 1 steps later
    (a result is valid until its lane's workspace is reused); returns the number of objects."""
    inflight, n_obj = [], 0
    for db, img in items:
        inflight.append(stage.run_device(db, img))
        if len(inflight) > stage.n_lanes - 1:
            n_obj += inflight.pop(0).n_obj
    for r in inflight:
        n_obj += r.n_obj
    stage.join()
    return n_obj


def job_sizes():
    from maze_image_processing_pipeline_b200.synth import synth_sizes
    return synth_sizes(SIZE_SEED, JOB_VIGNETTES, 64, 1024)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled every 20 ms DURING the timed region by a thread
    that calls NVML directly (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints; spawning nvidia-smi itself every few ms stalls the driver and was measured to slow the timed
    region by up to 2x).  Falls back to one nvidia-smi query before and after the region."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.h = None
        self.nv = None
        self._stop = threading.Event()
        self.t = None
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(reasons)))

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.02)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            bits = 0
            for bit, val in zip((0x8, 0x40, 0x20, 0x4), out[2:6]):
                if val.strip().lower().startswith("active"):
                    bits |= bit
            self.max_mhz = float(out[1])
            self.rows.append((time.time(), float(out[0]), bits))
        except Exception:
            pass

    def start(self):
        if self.nv is None:
            self._smi()
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self, t_begin, t_end):
        if self.nv is None:
            self._smi()
        else:
            self._stop.set()
            self.t.join(timeout=2)
        rows = self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source available"]}
        inside = [r for r in rows if t_begin <= r[0] <= t_end]
        how = "NVML, every 20 ms inside the timed region" if self.nv is not None else "nvidia-smi before/after the timed region"
        if not inside:
            mid = 0.5 * (t_begin + t_end)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:2]
            how += " (nearest samples: region shorter than the polling period)"
        bits = 0
        for r in inside:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": getattr(self, "max_mhz", None),
                "samples": len(inside), "source": how, "reasons": sorted(v for k, v in self.BITS.items() if bits & k)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's scipy restatement of the reference chain on host cores
# ------------------------------------------------------------------------------------------------
_CPU_MERGE = 0
_CPU_REF = False
_CPU_FULL = False


def _cpu_one(img):
    from oracle import scipy_chain
    try:
        mask, labels, table = scipy_chain.loki_chain(img, THRESHOLD, R_OPEN, R_CLOSE, merge_segments_distance=_CPU_MERGE,
                                                     use_reference=_CPU_REF)
    except TypeError:  # the reference's merge_labels raises when a bridge swallows a label (merge_labels.py:19-20)
        return (0, 0, None) if not _CPU_FULL else (0, 0, None, None, None)
    if _CPU_FULL:
        return int(labels.max()), int(mask.sum()), np.packbits(mask), labels.astype(np.int32), table
    return int(labels.max()), int(mask.sum()), None


def cpu_kind():
    """"reference": morphology + merge_labels are the reference's own files (oracle/_ref/, see oracle/make_ref.sh);
    "port": the restatement in oracle/scipy_chain.py."""
    from oracle import scipy_chain
    return "reference" if scipy_chain.reference_modules() is not None else "port"


def _cpu_pool(cores, merge=0, full=False):
    import multiprocessing as mp
    global _CPU_MERGE, _CPU_REF, _CPU_FULL
    _CPU_MERGE = merge  # inherited by the forked workers
    _CPU_REF = cpu_kind() == "reference"
    _CPU_FULL = full
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import oracle
    oracle.build()
    ctx = mp.get_context("fork")
    return ctx.Pool(cores)


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample_images(n, first=0):
    """The first n vignettes of the job from index `first`, regenerated on the host with the SAME
    sizes; pixel content comes from the numpy generator (same image model as the device generator)."""
    from maze_image_processing_pipeline_b200.synth import synth_vignette
    hs, ws = job_sizes()
    rng = np.random.default_rng(PIXEL_SEED + first)
    return [synth_vignette(rng, int(hs[(first + i) % JOB_VIGNETTES]), int(ws[(first + i) % JOB_VIGNETTES]))
            for i in range(n)]


def time_cpu(pool, imgs):
    t0 = time.perf_counter()
    out = pool.map(_cpu_one, imgs, chunksize=1)
    dt = time.perf_counter() - t0
    return dt, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_step = max(cores, min(4 * cores, 256))
    pool = _cpu_pool(cores, args.merge)
    imgs = cpu_sample_images(per_step)
    px = sum(int(i.size) for i in imgs)
    for _ in range(args.warmup):
        time_cpu(pool, imgs[: max(cores, per_step // 4)])
    total = 0.0
    for _ in range(args.steps):
        dt, _ = time_cpu(pool, imgs)
        total += dt
    pool.close()
    v = per_step * args.steps / total
    kind = cpu_kind()
    sample = (f"{per_step} vignettes ({px / 1e6:.1f} MPix) of configs[1] per step, multiprocessing.Pool({cores}); "
              + ("maze_ipp/isotropic.py + maze_ipp/merge_labels.py of the reference (oracle/_ref) around scipy.ndimage.label "
                 "and the C regionprops oracle" if kind == "reference" else "oracle/scipy_chain.py"))
    line = {
        "impl": "reference", "metric": "loki_vignettes_per_s", "value": v, "unit": "vignettes/s",
        "mpix_per_s": px * args.steps / total / 1e6, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, args.batch), "reference_sample_per_step": per_step,
        "cpu_baseline": {"value": v, "unit": "vignettes/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "vignettes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa(local_rank, world):
    """Multi-rank runs: pin this rank (and the packing threads it spawns) to the CPUs of the NUMA node its GPU
    hangs off, so that pinned staging buffers, the pack memcpy and the PCIe DMA stay on one socket."""
    if world <= 1 or os.environ.get("MAZE_NUMA_BIND", "1") == "0":
        return None
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def workload_config(args, batch):
    return {"workload": "configs[1]: 100k synthetic variable-size vignettes (64-1024 px, log-uniform), "
                        f"batches of {batch}", "batch_vignettes": batch, "threshold_brighter": THRESHOLD,
            "opening_radius": R_OPEN, "closing_radius": R_CLOSE, "merge_segments_distance": args.merge,
            "min_area": 0, "clear_border": False, "regionprops": "full table incl. high-order moments",
            "morphology": getattr(args, "morphology", "isotropic"),
            "e2e_batch_vignettes": getattr(args, "e2e_batch", batch),
            "l2": "each batch is larger than L2 (no flush needed)", "parallelism": f"images sharded over {args.gpus} GPU(s)",
            "pipeline": os.environ.get("MAZE_PIPELINE", "bands")}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from maze_image_processing_pipeline_b200 import _lib
    from maze_image_processing_pipeline_b200 import stage as S
    from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    cores_before = host_cores()
    numa = bind_to_gpu_numa(local_rank, world)
    if world > 1:
        # NCCL announces its version on stdout when NCCL_DEBUG asks for it: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        # share the host cores between the ranks for the packing threads
        os.environ.setdefault("MAZE_PACK_THREADS", str(max(1, min(8, cores_before // world))))
    _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hs, ws = job_sizes()
    B = args.batch
    n_batches_job = (JOB_VIGNETTES + B - 1) // B
    need = args.steps + args.warmup
    pp = S.SegmentationPostprocessingConfig(closing_radius=R_CLOSE, opening_radius=R_OPEN,
                                            merge_segments_distance=args.merge)
    stage = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(THRESHOLD), pp, merge_errors="ignore",
                                    morphology=args.morphology)

    # resident inputs: the batches this rank will touch, generated on the device.  At most MAX_RESIDENT distinct
    # batches stay in HBM (~0.5 GB of pixels each); longer runs cycle through them -- consecutive steps still work on
    # different batches, each larger than L2, and a batch comes round again only after ~12 GB of other pixels
    n_res = min(need, MAX_RESIDENT)
    batches = []
    for s in range(n_res):
        b = (rank + s * world) % n_batches_job
        lo = b * B
        hi = min(lo + B, JOB_VIGNETTES)
        geom = BatchGeometry(hs[lo:hi], ws[lo:hi])
        db = stage.prepare(DeviceBatch(geom))
        img = db.synth(PIXEL_SEED, lo)
        batches.append((db, img))
    stage.reserve([b[0].g for b in batches])
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def step(i):
        db, img = batches[i % n_res]
        res = stage.run_device(db, img)
        return res, res.mask

    for i in range(stage.n_lanes):  # every lane runs once (streams, workspaces, kernel attributes) before the warm-up
        step(i)[0].n_obj
    stage.join()
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record()
    n_vig = n_px = n_obj = 0
    inflight = []
    for i in range(args.warmup, need):
        res, _ = step(i)
        inflight.append(res)
        if len(inflight) > stage.n_lanes - 1:
            n_obj += inflight.pop(0).n_obj  # readback check n_lanes - 1 steps behind (rotating workspaces)
        n_vig += batches[i % n_res][0].g.n_img
        n_px += batches[i % n_res][0].g.pixels
    for r in inflight:
        n_obj += r.n_obj
    stage.join()
    ev1.record()
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # instrumented repeat of the same steps for the roofline: per-kernel CUDA-event durations, on ONE lane so that
    # an event pair brackets its own kernel only (with several lanes in flight the intervals of concurrent kernels
    # overlap and every duration is inflated by its neighbours)
    stage1 = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(THRESHOLD), pp, merge_errors="ignore",
                                     morphology=args.morphology, n_lanes=1)
    stage1.reserve([b[0].g for b in batches])
    for i in range(2):
        stage1.run_device(batches[i % n_res][0], batches[i % n_res][1]).n_obj
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record()
    for i in range(args.warmup, need):
        stage1.run_device(batches[i % n_res][0], batches[i % n_res][1])
    stage1.join()
    evp1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    ms_instr = evp0.elapsed_time(evp1)
    del stage1
    release()

    # end to end through the public stage call: host numpy in, host numpy out, copies timed
    # (host memory: every rank keeps its e2e inputs plus three pinned buffer sets; fewer batches per rank at N = 8)
    e2e_steps = max(1, min(args.steps, args.e2e_steps, max(6, 48 // world)))
    host_batches = []
    for i in range(args.warmup, args.warmup + e2e_steps):
        db, img = batches[i % n_res]
        # the end-to-end call uses batches of --e2e-batch vignettes (three pinned buffer sets per rank: smaller
        # batches keep the pinned working set of 8 ranks on one host in check; measured 2.4x faster at N = 8)
        n_k = min(db.g.n_img, args.e2e_batch)
        end = int(db.g.pix_off[n_k - 1]) + int(db.g.h[n_k - 1]) * int(db.g.w[n_k - 1])
        flat = img[:end].cpu().numpy()  # only the pixels of the vignettes the leg uses stay on the host
        host_batches.append([db.g.view(flat, k) for k in range(n_k)])
    def e2e_leg(st, materialize=False):
        """The streaming call a user makes -- host numpy in, results on the host, copies inside the timed region."""
        # warm-up: every lane allocates its device workspace on first use, map() rotates four pinned staging sets
        st.reserve([BatchGeometry.from_images(hb) for hb in host_batches[:2]])
        for r in st.map(host_batches[:min(len(host_batches), st.n_lanes + 2)]):
            pass
        barrier()
        t0 = time.perf_counter()
        nv = npx = up = down = 0
        for r in st.map(host_batches):
            if r.compact:
                down += r._runs.nbytes + r._band_out.nbytes + sum((0 if m is None else m.size) + (0 if l is None else 4 * l.size) for m, l in r._dense.values())
            else:
                down += r.geometry.total_px * 5
            if materialize:
                r.materialize()
            nv += len(r)
            npx += r.geometry.pixels
            up += r.geometry.total_px
            down += r.table.nbytes + r.lab_off.nbytes
        torch.cuda.synchronize()
        return time.perf_counter() - t0, nv, npx, up, down

    # (1) the headline: compact transport -- label images cross PCIe as run lists (8 B per run) and are expanded on
    #     demand (StageResult.labels(i) / object_mask); nothing dense is materialised inside the timed region
    stage_c = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(THRESHOLD), pp, merge_errors="ignore",
                                      morphology=args.morphology, compact=True)
    e2e_s, e2e_vig, e2e_px, h2d, d2h = e2e_leg(stage_c)
    # (2) contract-complete: every bool mask and int32 label image of the batch materialised on the host inside the
    #     timed region, (a) dense download as in round 1, (b) run lists + native multi-threaded expansion
    dense_s, dense_vig, _, _, dense_d2h = e2e_leg(stage)
    mat_s, mat_vig, _, _, _ = e2e_leg(stage_c, materialize=True)
    # resident throughput of the COMPACT step (what the end-to-end path runs on the device: no dense mask / label
    # image is written, the run list is the result) -- informational, `value` stays the dense step
    stage_c.reserve([b[0].g for b in batches])
    for i in range(stage_c.n_lanes):
        stage_c.run_device(batches[i % len(batches)][0], batches[i % len(batches)][1]).n_obj
    torch.cuda.synchronize()
    ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ec0.record()
    windowed_steps(stage_c, [batches[i % n_res] for i in range(args.warmup, need)])
    ec1.record()
    torch.cuda.synchronize()
    ms_compact = ec0.elapsed_time(ec1) / args.steps
    del stage_c
    release()

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_s, float(n_vig), float(n_px), float(e2e_vig), float(e2e_px), float(launches), dense_s,
                          mat_s, float(dense_vig), float(mat_vig)], dtype=torch.float64, device="cuda")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s, dense_s, mat_s = float(mx[0]), float(mx[1]), float(mx[7]), float(mx[8])
        n_vig, n_px, e2e_vig, e2e_px, launches, dense_vig, mat_vig = (float(sm[k]) for k in (2, 3, 4, 5, 6, 9, 10))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    px_per_launch = (n_px / world) / args.steps  # one rank's batch
    top = max(prof.items(), key=lambda kv: kv[1][0])
    top_name, (top_ms, top_cnt) = top
    bpp = KERNEL_BYTES_PER_PX.get(top_name, STAGE_BYTES_PER_PX)
    launches_per_step = top_cnt / args.steps
    alg_bytes = bpp * px_per_launch / max(launches_per_step, 1)
    avg_s = top_ms / top_cnt / 1e3
    achieved = alg_bytes / avg_s / 1e9
    traffic = measured_traffic(top_name)
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": None if traffic is None else traffic[0] * px_per_launch / max(launches_per_step, 1),
                "traffic_source": None if traffic is None else traffic[1], "peak_source": peak_src,
                "algorithmic_bytes_per_px": bpp, "avg_launch_ms": top_ms / top_cnt,
                "share_of_step": top_ms / ms_instr,
                "timing": "CUDA events around every launch of an instrumented single-lane repeat of the timed steps"}
    stage_gbs = STAGE_BYTES_PER_PX * n_px / (ms / 1e3) / 1e9
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}

    # ---- parity gate + CPU baseline: the oracle chain on the first vignettes of the first timed batch (the SAME pixels
    # the GPU processed) -- timed on all host cores (N = 1 only), and compared with the GPU's masks / label images /
    # object tables in both transports.  No value is printed when the comparison fails.
    cpu_baseline = parity = None
    if not args.no_cpu_baseline:
        cores = host_cores()
        n_s = max(cores, min(4 * cores, 256)) if world == 1 else 16
        pool = _cpu_pool(cores, args.merge, full=True)
        imgs = host_batches[0][:n_s]
        if world == 1:
            time_cpu(pool, imgs[: max(1, len(imgs) // 4)])
        dt, want = time_cpu(pool, imgs)
        pool.close()
        if world == 1:
            kind = cpu_kind()
            cpu_baseline = {"value": len(imgs) / dt, "unit": "vignettes/s", "cores": cores, "kind": kind,
                            "mpix_per_s": sum(int(i.size) for i in imgs) / dt / 1e6,
                            "sample": f"first {len(imgs)} vignettes of the first timed batch (same pixels as the GPU run), "
                                      + ("the reference's own isotropic.py / merge_labels.py (oracle/_ref) around "
                                         "scipy.ndimage.label + the C regionprops oracle" if kind == "reference"
                                         else "oracle/scipy_chain.py (the reference chain on scipy.ndimage)")
                                      + f", Pool({cores})"}
        import oracle as _oracle
        masks_eq = labels_eq = True
        max_rel = 0.0
        n_cmp = 0
        cols = list(range(1, 8)) + list(range(_oracle.F_HU, _oracle.F_HU + 7)) + [49, 50, 51, 53, 54, 55, 56]
        for compact in (False, True):
            st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(THRESHOLD), pp, merge_errors="ignore",
                                         morphology=args.morphology, compact=compact)
            got = st(imgs)
            failed = set() if got.merge_failed is None else set(int(i) for i in got.merge_failed)
            for i, (im, w) in enumerate(zip(imgs, want)):
                if w[2] is None or i in failed:  # merge_labels raised in the reference
                    continue
                n_cmp += 1
                masks_eq &= bool(np.array_equal(np.packbits(got.mask(i)), w[2]))
                labels_eq &= bool(np.array_equal(got.labels(i), w[3]))
                f, t = got.features(i), w[4]
                k = min(len(f), len(t))
                sel = t[:k, _oracle.F_AREA] > 0
                a_, b_ = f[:k][sel][:, cols], t[:k][sel][:, cols]
                if a_.size:
                    max_rel = max(max_rel, float(np.nanmax(np.abs(a_ - b_) / np.maximum(np.abs(b_), 1e-9))))
                labels_eq &= bool(np.array_equal(f[:k, _oracle.F_AREA], t[:k, _oracle.F_AREA]))
            del st, got
            release()
        parity = {"vignettes": n_cmp // 2, "transports": ["dense", "compact"], "masks_equal": masks_eq,
                  "labels_equal": labels_eq, "table_max_rel": max_rel, "table_rel_tolerance": 1e-5,
                  "checked_against": cpu_kind()}
        if not (masks_eq and labels_eq and max_rel <= 1e-5):
            print(json.dumps({"error": "parity gate failed: no value is reported", "parity": parity}))
            if world > 1:
                dist.destroy_process_group()
            return 2

    # ---- variants SURVEY.md 8d asks to report next to the headline: resident throughput with merge_labels on
    # (merge_segments_distance = 10), with the live pipeline's footprints (morphology = crosses) and with the label
    # filters on (clear_border + min_area = 12: applied on the run list inside the band pipeline)
    variants = None
    if world == 1 and not args.no_variants and args.merge == 0 and args.morphology == "isotropic":
        variants = {}
        vb = batches[args.warmup % n_res][0], batches[args.warmup % n_res][1]
        for name, kw, vsteps in (("merge10", dict(merge_segments_distance=10), 8), ("crosses", dict(), 8),
                                 ("label_filters", dict(clear_border=True, min_area=12), 8),
                                 ("threshold_branch", None, 8)):
            # (threshold_branch: the reference's shipped `threshold` segmentation, loki/pipeline.py:648-656 -- mask +
            # one region per vignette, no morphology / labelling)
            vpp = None if kw is None else S.SegmentationPostprocessingConfig(closing_radius=R_CLOSE, opening_radius=R_OPEN, **kw)
            st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(THRESHOLD), vpp, merge_errors="ignore",
                                         morphology="crosses" if name == "crosses" else "isotropic")
            st.reserve([vb[0].g])
            for _ in range(2 * st.n_lanes):  # every lane allocates (and then grows) its workspace on its first uses
                st.run_device(*vb).n_obj
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            windowed_steps(st, [vb] * vsteps)
            e1.record()
            torch.cuda.synchronize()
            vms = e0.elapsed_time(e1) / vsteps
            variants[name] = {"value": vb[0].g.n_img / (vms / 1e3), "unit": "vignettes/s", "ms_per_step": vms,
                              "steps": vsteps, "note": "resident, one 4096-vignette batch repeated"}
            del st
            release()

    line = {
        "metric": "loki_vignettes_per_s", "value": n_vig / (ms / 1e3), "unit": "vignettes/s",
        "mpix_per_s": n_px / (ms / 1e3) / 1e6, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, B),
        "objects_per_step": n_obj / args.steps,
        "roofline": roofline,
        "stage_hbm": {"achieved": stage_gbs / world, "unit": "GB/s per GPU", "bytes_per_px": STAGE_BYTES_PER_PX,
                      "frac_of_measured": stage_gbs / world / peak, "frac_of_nominal_8TBps": stage_gbs / world / 8000.0},
        "kernels": kernels, "ms_per_step_instrumented": ms_instr / args.steps,
        "resident_compact": {"ms_per_step": ms_compact, "value": (n_vig / world / args.steps) / (ms_compact / 1e3) * world,
                             "unit": "vignettes/s", "note": "the step of the end-to-end path: bit plane + run list + "
                             "object table, no dense mask / label image (rank 0's time)"},
        "cpu_baseline": cpu_baseline, "parity": parity, "variants": variants,
        "e2e": {"value": e2e_vig / e2e_s, "unit": "vignettes/s", "mpix_per_s": e2e_px / e2e_s / 1e6,
                "steps": e2e_steps, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                "transport": "compact: label images cross PCIe as run lists {y, x0, x1, label} (8 B per run) + object "
                             "table; StageResult.mask(i) / labels(i) / object_mask expand them on the host on demand"},
        "e2e_materialized": {
            "note": "contract-complete: every bool mask and int32 label image of the batch is a host array inside the "
                    "timed region",
            "dense_download": {"value": dense_vig / dense_s, "unit": "vignettes/s",
                               "d2h_bytes_per_step": dense_d2h // e2e_steps},
            "run_list_expand": {"value": mat_vig / mat_s, "unit": "vignettes/s", "d2h_bytes_per_step": d2h // e2e_steps}},
        "gpu_launches": int(launches), "clocks": clocks,
        "device_memory_peak_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--merge", type=int, default=0, help="merge_segments_distance (0 = off, the schema default)")
    ap.add_argument("--morphology", default="isotropic", choices=["isotropic", "crosses"],
                    help="isotropic = maze_ipp/isotropic.py (north star); crosses = the live pipeline's "
                         "binary_opening / binary_closing with disk(r, decomposition='crosses')")
    ap.add_argument("--e2e-steps", type=int, default=16,
                    help="batches of the streaming end-to-end leg (a 100k-vignette job is 49 batches of 2048)")
    ap.add_argument("--e2e-batch", type=int, default=2048, help="vignettes per stage.map batch in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline AND the parity gate")
    ap.add_argument("--no-variants", action="store_true", help="skip the merge10 / crosses variant block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
