"""Debug helper: vignettes of the bench batch on which the windowed and the whole-image merge kernels disagree."""
import sys, os
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
labels = res.labels.clone()
outs = {}
for flag in ("0", "1"):
    os.environ["MAZE_MERGE_WINDOWED"] = flag
    lab = labels.clone()
    out = db.merge_labels(lab, lab, res.lab_off, int(res.lab_off[-1]), 10.0)
    torch.cuda.synchronize()
    outs[flag] = (lab.cpu().numpy(), out[1].cpu().numpy(), out[3].cpu().numpy(), out[2].cpu().numpy())
bad = np.nonzero(outs["0"][2] != outs["1"][2])[0]
print("status differs on", bad)
host = labels.cpu().numpy()
save = {}
for i in bad[:8]:
    save[f"in_{i}"] = g.view(host, i).copy()
    save[f"old_{i}"] = g.view(outs["0"][0], i).copy()
    save[f"new_{i}"] = g.view(outs["1"][0], i).copy()
    print(i, "old status", outs["0"][2][i], "nm", outs["0"][1][i], "idx state", outs["0"][3][2 * i:2 * i + 2],
          "| new status", outs["1"][2][i], "nm", outs["1"][1][i], "idx state", outs["1"][3][2 * i:2 * i + 2])
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/merge_dbg.npz", **save)
