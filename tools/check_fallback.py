import sys, os, time
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
for b in range(8):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = st.prepare(DeviceBatch(g))
    batches.append((db, db.synth(bench.PIXEL_SEED, b * B)))
st.reserve([b[0].g for b in batches])
for db, img in batches:
    r = st.run_device(db, img)
    n = db.g.n_img
    torch.cuda.synchronize()
    ws_ = st._ws_ring[(st._ws_i - 1) % 2]
    c = ws_._t["counts"][:3 * n].cpu().numpy()
    left = db.fused_lists()[2]
    print("fallback flags", int((c[n:2 * n] != 0).sum()), "acc_base<0 & nlab>0", int(((c[2 * n:] < 0) & (c[:n] > 0)).sum()), "left", len(left), "n_obj", r.n_obj)
for mode in ("nofinalize", "finalize_prev"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    prev = None
    for rep in range(5):
        for db, img in batches:
            r = st.run_device(db, img)
            if mode == "finalize_prev" and prev is not None:
                prev.n_obj
            prev = r
    torch.cuda.synchronize()
    print(mode, (time.perf_counter() - t0) / 40 * 1e3, "ms/step")
