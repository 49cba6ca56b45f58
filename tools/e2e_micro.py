import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n = 250_000_000
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
lab = torch.empty(n, dtype=torch.int32, device="cuda")
hl = torch.empty(n, dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize()
sl = pin[:n - 1000]
print("slice pinned:", sl.is_pinned())
t = time.perf_counter(); d2 = sl.to("cuda", non_blocking=True); t1 = time.perf_counter() - t; torch.cuda.synchronize(); print(".to() call ms", t1 * 1e3, "total", (time.perf_counter() - t) * 1e3)
t = time.perf_counter(); d2 = sl.to("cuda", non_blocking=True); t1 = time.perf_counter() - t; torch.cuda.synchronize(); print(".to() 2nd call ms", t1 * 1e3, "total", (time.perf_counter() - t) * 1e3)
t = time.perf_counter(); dev[:n - 1000].copy_(sl, non_blocking=True); t1 = time.perf_counter() - t; torch.cuda.synchronize(); print("copy_ call ms", t1 * 1e3, "total", (time.perf_counter() - t) * 1e3)
cs = torch.cuda.Stream()
t = time.perf_counter()
with torch.cuda.stream(cs):
    hl.copy_(lab, non_blocking=True)
t1 = time.perf_counter() - t
# host work meanwhile
a = np.empty(n, np.uint8); b = pin.numpy()
t2 = time.perf_counter(); a[:] = b; t3 = time.perf_counter() - t2
cs.synchronize(); print("D2H enqueue ms", t1 * 1e3, "host memcpy ms", t3 * 1e3, "total", (time.perf_counter() - t) * 1e3)
