import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200._lib import FUSED_CAPS
hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
for b in range(0, 48):
    g = BatchGeometry(hs[b * B:(b + 1) * B], ws[b * B:(b + 1) * B])
    db = st.prepare(DeviceBatch(g))
    img = db.synth(bench.PIXEL_SEED, b * B)
    r = st.run_device(db, img)
    torch.cuda.synchronize()
    n = g.n_img
    ws_ = st._ws_ring[(st._ws_i - 1) % st.n_lanes]
    c = ws_._t["counts"][:3 * n].cpu().numpy()
    bad = np.nonzero(c[n:2 * n])[0]
    for i in bad:
        h, w = int(g.h[i]), int(g.w[i]); words = int(g.nwords[i])
        cap = [x for x in FUSED_CAPS if words <= x][0]
        RC = (4 * (2 * cap - words) - 2 * (h + 2)) // 10
        m = g.view(img.cpu().numpy(), i) > 40
        runs = int((np.diff(np.pad(m.astype(np.int8), ((0, 0), (1, 0))), axis=1) == 1).sum())
        print("batch", b, "img", i, "h,w", h, w, "words", words, "cap", cap, "RC", RC, "threshold-mask runs", runs, "fg frac", m.mean())
    r.n_obj
print("done")
