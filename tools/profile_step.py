"""Host-side profile of one resident step (where does the non-kernel time go?)."""
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
batches = []
for b in range(4):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = DeviceBatch(g)
    batches.append((db, db.synth(1, b * B)))
for db, img in batches:
    st.run_device(db, img)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    for db, img in batches:
        st.run_device(db, img)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 12 * 1e3)
pr = cProfile.Profile()
pr.enable()
for db, img in batches:
    st.run_device(db, img)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
