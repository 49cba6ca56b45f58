"""PCIe ceiling of the box for the end-to-end leg: D2H of the label image + mask of one 2048-vignette batch
(1.0 GB + 0.25 GB) alone, together with the H2D of the next batch, and with the host packing threads running."""
import json, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import torch.distributed as dist
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
bench.bind_to_gpu_numa(LOCAL, WORLD)
if WORLD > 1:  # all ranks copy at the same time: the aggregate ceiling of the host (torchrun --nproc-per-node N)
    dist.init_process_group("gloo")
n = 245_000_000
d_lab = torch.empty(n, dtype=torch.int32, device="cuda"); d_mask = torch.empty(n, dtype=torch.uint8, device="cuda")
d_img = torch.empty(n, dtype=torch.uint8, device="cuda")
h_lab = torch.empty(n, dtype=torch.int32, pin_memory=True); h_mask = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_img = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, pack, reps=5):
    best = 1e9
    src = np.empty(n, np.uint8)
    for _ in range(reps):
        torch.cuda.synchronize()
        if WORLD > 1:
            dist.barrier()
        stop = []
        th = None
        if pack:
            def work():
                dst = h_img.numpy()
                while not stop:
                    dst[:] = src
            th = threading.Thread(target=work); th.start()
        t = time.perf_counter()
        with torch.cuda.stream(s1):
            h_lab.copy_(d_lab, non_blocking=True); h_mask.copy_(d_mask, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2):
                d_img.copy_(h_img, non_blocking=True)
        s1.synchronize()
        dt = time.perf_counter() - t
        stop.append(1)
        if th: th.join()
        torch.cuda.synchronize()
        best = min(best, dt)
    return best


for h2d, pack in ((False, False), (True, False), (True, True)):
    dt = run(h2d, pack)
    if WORLD > 1:
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        if dist.get_rank() != 0:
            continue
    print(json.dumps({"ranks": WORLD, "aggregate_d2h_GBps": round(WORLD * 5 * n / dt / 1e9, 1),"d2h_GBps": round(5 * n / dt / 1e9, 1), "ms": round(dt * 1e3, 2), "with_h2d": h2d, "with_host_memcpy": pack,
                      "vignettes_per_s_ceiling": round(WORLD * 2048 / dt)}))
