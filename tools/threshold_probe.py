import sys, time, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S, _lib
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
g = BatchGeometry(hs[:4096], ws[:4096]); db = DeviceBatch(g); img = db.synth(1, 0)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40))
st.reserve([g])
for _ in range(st.n_lanes):
    r = st.run_device(db, img); r.n_obj
torch.cuda.synchronize()
_lib.prof_enable(True)
t0 = time.perf_counter()
rs = [st.run_device(db, img) for _ in range(4)]
for r in rs: r.n_obj
st.join(); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 4
_lib.prof_enable(False)
prof = _lib.prof_collect()
print("threshold branch ms/step", dt * 1e3, "redone", [r.redone for r in rs], "dense_only", len(rs[0].dense_only), "runs", rs[0].n_runs)
print({k: round(v[0] / 4, 3) for k, v in prof.items()})
# the bench's variant measurement, step by step
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    bench.windowed_steps(st, [(db, img)] * 8)
    e1.record()
    torch.cuda.synchronize()
    print("windowed x8: event ms/step", e0.elapsed_time(e1) / 8, "wall ms/step", (time.perf_counter() - t0) / 8 * 1e3)
