"""Per-step GPU and host time of the resident loop, to find sporadic hiccups."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
batches = []
for b in range(43):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = st.prepare(DeviceBatch(g))
    batches.append((db, db.synth(bench.PIXEL_SEED, b * B)))
st.reserve([b[0].g for b in batches])
for db, img in batches[:3]:
    st.run_device(db, img)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
host = []
prev = None
evs[0].record()
for i, (db, img) in enumerate(batches[3:]):
    t0 = time.perf_counter()
    r = st.run_device(db, img)
    t1 = time.perf_counter()
    if prev is not None:
        prev.n_obj
    t2 = time.perf_counter()
    prev = r
    evs[i + 1].record()
    host.append((t1 - t0, t2 - t1))
torch.cuda.synchronize()
gpu = [evs[i].elapsed_time(evs[i + 1]) for i in range(40)]
print("total ms/step", sum(gpu) / 40)
print("gpu per step:", " ".join(f"{x:.2f}" for x in gpu))
print("host enqueue:", " ".join(f"{x[0]*1e3:.2f}" for x in host))
print("host finalize:", " ".join(f"{x[1]*1e3:.2f}" for x in host))
print("gc counts", gc.get_count())
for w_ in st._ws_ring:
    print("arena buf", w_.arena.buf.numel() / 1e6, "MB need(last)", w_.arena.need / 1e6)
print("left per batch:", [len(b[0].fused_lists()[2]) for b in batches[3:]])
print("px per batch (M):", [round(b[0].g.total_px / 1e6) for b in batches[3:]])
