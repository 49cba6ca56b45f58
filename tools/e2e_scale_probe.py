"""End-to-end leg only (stage.map, compact transport), under torchrun or alone: vignettes/s over all ranks for a
choice of packing threads / NUMA binding / prefetch depth (env: MAZE_PACK_THREADS, MAZE_NUMA_BIND).  Used to find the
host-side limit of the multi-GPU end-to-end path."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch

rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
cores = bench.host_cores()
node = bench.bind_to_gpu_numa(lr, world)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    os.environ.setdefault("MAZE_PACK_THREADS", str(max(1, min(8, cores // world))))
hs, ws = bench.job_sizes()
B, NB = 2048, int(os.environ.get("NB", "12"))
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=True)
hb = []
for b in range(NB):
    lo = (rank * NB + b) * B
    g = BatchGeometry(hs[lo:lo + B], ws[lo:lo + B])
    flat = DeviceBatch(g).synth(1, lo).cpu().numpy()
    hb.append([g.view(flat, k) for k in range(g.n_img)])
st.reserve([BatchGeometry.from_images(x) for x in hb[:2]])
for r in st.map(hb[:8]):
    pass
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
for r in st.map(hb):
    n += len(r)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt, float(n)], dtype=torch.float64, device="cuda")
if world > 1:
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    dt, n = float(mx[0]), float(sm[1])
if rank == 0:
    print(json.dumps({"world": world, "cores": cores, "affinity": len(os.sched_getaffinity(0)), "numa_node": node,
                      "pack_threads": os.environ.get("MAZE_PACK_THREADS"), "e2e_vignettes_per_s": n / dt,
                      "ms_per_batch_per_rank": dt / NB * 1e3}))
if world > 1:
    dist.destroy_process_group()
