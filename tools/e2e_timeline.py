import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=bool(int(os.environ.get("COMPACT", "1"))))
hb = []
for b in range(8):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = DeviceBatch(g)
    flat = db.synth(1, b * B).cpu().numpy()
    hb.append([g.view(flat, k) for k in range(g.n_img)])
for r in st.map(hb[:3]):
    pass
T = {"enq": [], "cmp": []}
oe, oc = st._enqueue, st._complete
def enq(*a, **k):
    t = time.perf_counter(); r = oe(*a, **k); T["enq"].append(time.perf_counter() - t); return r
def cmp(*a, **k):
    t = time.perf_counter(); r = oc(*a, **k); T["cmp"].append(time.perf_counter() - t); return r
st._enqueue, st._complete = enq, cmp
t0 = time.perf_counter()
for r in st.map(hb):
    pass
print("ms/batch", (time.perf_counter() - t0) / len(hb) * 1e3)
print("enqueue ms:", " ".join(f"{x*1e3:.1f}" for x in T["enq"]))
print("complete ms:", " ".join(f"{x*1e3:.1f}" for x in T["cmp"]))
# inside enqueue
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for r in st.map(hb[:4]):
    pass
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(10)
