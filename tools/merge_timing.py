"""Per-vignette phase times of k_merge_windowed (needs the -DMW_TIMING build: sh csrc/build.sh -DMW_TIMING -o
csrc/libmaze_b200_timing.so).  Swaps the library in before the package loads it."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from maze_image_processing_pipeline_b200 import _lib
_lib.SO_PATH = os.path.join(_lib.CSRC, "libmaze_b200_timing.so")
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 4096
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
g = BatchGeometry(hs[:B], ws[:B]); db = DeviceBatch(g)
img = db.synth(bench.PIXEL_SEED, 0)
res = st.run_device(db, img).finalize()
tab = res._table[:, :6].cpu().numpy(); off = res.lab_off.cpu().numpy().astype(np.int64)
need = S._merge_candidates(tab, off, g.h, g.w, 10.0)
print("candidates", len(need))
h = _lib.lib()
h.maze_merge_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
for rep in range(2):
    h.maze_merge_debug_clear()
    lab = res.labels.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = db.merge_labels(lab, lab, res.lab_off, int(off[-1]), 10.0, only=need)
    e1.record(); torch.cuda.synchronize()
print("merge call ms", e0.elapsed_time(e1))
os.environ["MAZE_MERGE_WINDOWED"] = "0"
lab0 = res.labels.clone()
out0 = db.merge_labels(lab0, lab0, res.lab_off, int(off[-1]), 10.0, only=need)
torch.cuda.synchronize()
os.environ["MAZE_MERGE_WINDOWED"] = "1"
print("same as the whole-image kernel:", bool(torch.equal(lab, lab0)), bool(torch.equal(out[0], out0[0])),
      bool(torch.equal(out[1], out0[1])), bool(torch.equal(out[3], out0[3])))
dbg = np.zeros((B, 16), np.int64)
h.maze_merge_debug_read(dbg.ctypes.data, B)
d = dbg[need].astype(np.float64) / 1e3  # us
names = ["total", "tables", "edt0", "mins0", "dirty", "pop", "edt", "summin", "fill", "outside", "_", "_", "horiz", "envelope"]
print("sum over candidates (ms):", {n: round(float(d[:, k].sum()) / 1e3, 2) for k, n in enumerate(names) if n != "_"})
print("CTA-time / (148 SMs x 2):", round(d[:, 0].sum() / 1e3 / 296, 3), "ms; longest vignette", round(d[:, 0].max() / 1e3, 3), "ms")
order = np.argsort(-d[:, 0])[:12]
for j in order:
    i = need[j]
    print(f"vig {i:5d} px {g.npx[i]:8d} labels {off[i+1]-off[i]:3d} pops {int(dbg[i,10])-1:3d} cta {int(dbg[i,11]):4d} | " +
          " ".join(f"{n}={d[j,k]:.0f}" for k, n in enumerate(names) if n != "_"))
