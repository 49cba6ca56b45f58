"""Host-side profile of the end-to-end stage call (numpy in, numpy out)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch

hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
hb = []
for b in range(3):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = DeviceBatch(g)
    flat = db.synth(1, b * B).cpu().numpy()
    hb.append([g.view(flat, k) for k in range(g.n_img)])
for h in hb:
    st(h)
torch.cuda.synchronize()
t0 = time.perf_counter()
for h in hb:
    st(h)
print("ms/batch", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile()
pr.enable()
for h in hb:
    st(h)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
