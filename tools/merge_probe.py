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
g = BatchGeometry(hs[:B], ws[:B]); db = DeviceBatch(g)
img = db.synth(bench.PIXEL_SEED, 0)
res = st.run_device(db, img).finalize()
nlab = (res.lab_off[1:] - res.lab_off[:-1]).cpu().numpy()
print("labels per vignette: mean", nlab.mean(), "max", nlab.max(), "hist", np.bincount(np.minimum(nlab, 12)))
labels = res.labels.clone()
def run(sel):
    gg = g.subset(sel); sb = DeviceBatch(gg)
    off = np.concatenate([[0], np.cumsum(nlab[sel])]).astype(np.int32)
    d_off = torch.from_numpy(off).cuda(); nobj = int(off[-1])
    lab = labels.clone()
    torch.cuda.synchronize(); t = time.perf_counter()
    out = sb.merge_labels(lab, lab, d_off, nobj, 10.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    return dt, out[1].cpu().numpy(), out[3].cpu().numpy()
allsel = np.arange(B)
def run_labels(windowed):
    os.environ["MAZE_MERGE_WINDOWED"] = "1" if windowed else "0"
    lab = labels.clone()
    out = db.merge_labels(lab, lab, res.lab_off, int(res.lab_off[-1]), 10.0)
    torch.cuda.synchronize()
    return lab, out
la, oa = run_labels(False)
lb, ob = run_labels(True)
print("windowed == whole-image kernel: labels", bool(torch.equal(la, lb)), "dists", bool(torch.equal(oa[0], ob[0])),
      "n_merge", bool(torch.equal(oa[1], ob[1])), "status", bool(torch.equal(oa[3], ob[3])))
for wflag in ("0", "1"):
    os.environ["MAZE_MERGE_WINDOWED"] = wflag
    run(allsel)
    dt, nm, stt = run(allsel); print("windowed", wflag, "all: ms", dt, "merges", nm.sum(), "errors", (stt != 0).sum(),
                                     "merge hist", np.bincount(np.minimum(nm, 12)))
px = g.npx
for lo, hi in [(0, 2e4), (2e4, 1e5), (1e5, 3e5), (3e5, 2e6)]:
    sel = np.nonzero((px >= lo) & (px < hi))[0]
    dt, nm, stt = run(sel); print(f"px in [{lo:.0f},{hi:.0f}): n={len(sel)} ms={dt:.2f} labels mean={nlab[sel].mean():.2f}")
# the slowest single vignettes
big = np.argsort(-px)[:4]
for i in big:
    dt, nm, stt = run(np.array([i])); print("vignette", i, "px", px[i], "labels", nlab[i], "ms", round(dt, 3), "merges", nm.sum())
many = np.argsort(-nlab)[:4]
for i in many:
    dt, nm, stt = run(np.array([i])); print("vignette", i, "px", px[i], "labels", nlab[i], "ms", round(dt, 3), "merges", nm.sum())
