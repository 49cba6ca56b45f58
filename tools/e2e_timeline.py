"""Per-batch timeline of the streaming end-to-end call (stage.map): time between yields, time inside _prepack /
_enqueue / _complete.  COMPACT=0/1 selects the transport, NB the number of batches, PACK_THREADS the copy threads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 2048
NB = int(os.environ.get("NB", "16"))
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=bool(int(os.environ.get("COMPACT", "1"))))
hb = []
for b in range(NB):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = DeviceBatch(g)
    flat = db.synth(1, b * B).cpu().numpy()
    hb.append([g.view(flat, k) for k in range(g.n_img)])
for r in st.map(hb[:5]):
    pass
T = {"pre": [], "enq": [], "cmp": []}
for name, key in (("_prepack", "pre"), ("_enqueue", "enq"), ("_complete", "cmp")):
    orig = getattr(st, name)
    def wrap(*a, _o=orig, _k=key, **k):
        t = time.perf_counter(); r = _o(*a, **k); T[_k].append(time.perf_counter() - t); return r
    setattr(st, name, wrap)
t0 = time.perf_counter()
stamps = []
for r in st.map(hb):
    stamps.append(time.perf_counter() - t0)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("ms/batch", tot / len(hb) * 1e3, "=>", len(hb) * B / tot, "vignettes/s")
print("yield gaps ms:", " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip([0] + stamps[:-1], stamps)))
for k in T:
    print(k, "ms:", " ".join(f"{x * 1e3:.1f}" for x in T[k]))
