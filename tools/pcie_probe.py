"""PCIe ceiling of the box for the end-to-end leg: D2H of the label image + mask of one 2048-vignette batch
(1.0 GB + 0.25 GB) alone, together with the H2D of the next batch, and with the host packing threads running."""
import json, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
bench.bind_to_gpu_numa(0, 1)
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
    print(json.dumps({"d2h_GBps": round(5 * n / dt / 1e9, 1), "ms": round(dt * 1e3, 2), "with_h2d": h2d, "with_host_memcpy": pack,
                      "vignettes_per_s_ceiling": round(2048 / dt)}))
