"""Host time of one run_device() call on a single 4096 x 4096 frame vs. the GPU time of the step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_dense_frame
frame = synth_dense_frame(11, size=4096, n_blobs=3000)
g = BatchGeometry([4096], [4096]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
for compact in (False, True):
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=compact)
    db = st.prepare(DeviceBatch(g)); st.reserve([g])
    for _ in range(8):
        st.run_device(db, d).n_obj
    torch.cuda.synchronize()
    N = 60
    t0 = time.perf_counter(); host = 0.0
    rs = []
    for i in range(N):
        t = time.perf_counter()
        rs.append(st.run_device(db, d))
        host += time.perf_counter() - t
        if len(rs) >= st.n_lanes - 1:
            rs.pop(0).n_obj
    for r in rs:
        r.n_obj
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"compact={compact}: wall {wall / N * 1e3:.3f} ms per frame, host time inside run_device {host / N * 1e3:.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(30):
    st.run_device(db, d).n_obj
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
