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
for b in range(6):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = st.prepare(DeviceBatch(g))
    batches.append((db, db.synth(1, b * B)))
for db, img in batches:
    st.run_device(db, img)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    for db, img in batches:
        st.run_device(db, img)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host enqueue ms/step", t_host / 18 * 1e3, "total ms/step", t_all / 18 * 1e3)
pr = cProfile.Profile()
pr.enable()
for db, img in batches:
    st.run_device(db, img)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
