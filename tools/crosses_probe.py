import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S, _lib
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
g = BatchGeometry(hs[:4096], ws[:4096]); db = DeviceBatch(g); img = db.synth(1, 0)
for morph in ("isotropic", "crosses"):
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, morphology=morph)
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
    print(morph, "ms/step", dt * 1e3, "redone", [r.redone for r in rs], "dense_only", len(rs[0].dense_only), st._passes())
    print({k: round(v[0] / 4, 3) for k, v in prof.items()})
