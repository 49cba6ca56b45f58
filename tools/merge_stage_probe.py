"""Where the time of a step with merge_segments_distance = 10 goes: per-kernel CUDA-event times (maze_prof) and the wall
clock of the host-side pieces, on the bench's 4096-vignette batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S, _lib
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 4096
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, merge_segments_distance=10)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, merge_errors="ignore")
g = BatchGeometry(hs[:B], ws[:B]); db = DeviceBatch(g)
img = db.synth(bench.PIXEL_SEED, 0)
st.reserve([g])
for _ in range(4):
    st.run_device(db, img).n_obj
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    r = st.run_device(db, img); r.n_obj
torch.cuda.synchronize()
print("ms per step (sync each)", (time.perf_counter() - t) / 5 * 1e3)
_lib.prof_enable(True)
for _ in range(3):
    st.run_device(db, img).n_obj
torch.cuda.synchronize()
_lib.prof_enable(False)
for k, v in sorted(_lib.prof_collect().items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {v[0] / 3:8.3f} ms per step  {v[1] // 3} launches")
# host pieces
res = st._run_fused_async(db, img, img, 40, st._passes()) if False else None
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    st.run_device(db, img).n_obj
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
