"""How many vignettes of the WHOLE configs[1] job (100k vignettes, 25 batches of 4096) leave the band pipeline for
the per-operator kernels (oversize for a band, run tables full, phantom pixel, staging rows exhausted), and how many
runs / objects a batch holds against the capacities of its buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 4096
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=bool(int(os.environ.get("COMPACT", "0"))))
tot_dense = tot_redo = 0
for b in range((bench.JOB_VIGNETTES + B - 1) // B):
    lo, hi = b * B, min((b + 1) * B, bench.JOB_VIGNETTES)
    g = BatchGeometry(hs[lo:hi], ws[lo:hi])
    db = st.prepare(DeviceBatch(g))
    img = db.synth(bench.PIXEL_SEED, lo)
    r = st.run_device(db, img)
    n_obj = r.n_obj
    torch.cuda.synchronize()
    tot_dense += len(r.dense_only); tot_redo += int(r.redone)
    print(f"batch {b:2d}: vignettes {g.n_img} MPix {g.pixels / 1e6:.0f} runs {r.n_runs} (cap {max(g.total_words // 3, 1 << 16)}) "
          f"objects {n_obj} (cap {S._stage_cap(g)}) per-operator vignettes {len(r.dense_only)} redone {r.redone}")
print("job: vignettes without a run list", tot_dense, "batches redone", tot_redo)
