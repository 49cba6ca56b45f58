"""Host-side timing of the end-to-end streaming call (numpy in, numpy out)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch

hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=bool(int(os.environ.get("COMPACT", "1"))))
hb = []
for b in range(6):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = DeviceBatch(g)
    flat = db.synth(1, b * B).cpu().numpy()
    hb.append([g.view(flat, k) for k in range(g.n_img)])
for r in st.map(hb[:3]):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in st.map(hb):
    pass
print("map ms/batch", (time.perf_counter() - t0) / len(hb) * 1e3)
# pieces
g = BatchGeometry.from_images(hb[0])
pin = torch.empty(g.total_px, dtype=torch.uint8, pin_memory=True)
for th in (1, 4, 8, 16):
    t0 = time.perf_counter(); g.pack_host(hb[0], out=pin.numpy(), threads=th); print("pack threads", th, (time.perf_counter() - t0) * 1e3, "ms")
t0 = time.perf_counter(); BatchGeometry.from_images(hb[0]); print("geometry ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); db = DeviceBatch(g); torch.cuda.synchronize(); print("DeviceBatch ms", (time.perf_counter() - t0) * 1e3)
d = torch.empty(g.total_px, dtype=torch.int32, device="cuda"); hp = torch.empty(g.total_px, dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); hp.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H labels", dt * 1e3, "ms", g.total_px * 4 / dt / 1e9, "GB/s")
d8 = torch.empty(g.total_px, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d8.copy_(pin, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D image", dt * 1e3, "ms", g.total_px / dt / 1e9, "GB/s")
pr = cProfile.Profile(); pr.enable()
for r in st.map(hb):
    pass
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
