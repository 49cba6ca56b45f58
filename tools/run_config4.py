"""BASELINE.json configs[4]: 1 000 000 synthetic vignettes (sizes as configs[1], seed 2) sharded by image over the
ranks (no collective on the data path).  Every batch of 4096 is generated on the device, run through the stage
(dense outputs) and checked with size-independent properties on the device: the areas of the object table add up to
the foreground pixel count, labels are exactly the non-zero pixels of the mask, the largest label of every vignette
equals its label count.  Timed: the stage steps only (CUDA events; generation and checks are outside).

    python tools/run_config4.py [--vignettes N]           # one GPU
    torchrun --nproc-per-node 8 tools/run_config4.py      # eight (each rank takes every 8th batch)
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_sizes

ap = argparse.ArgumentParser()
ap.add_argument("--vignettes", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--check-every", type=int, default=8, help="property checks on every k-th batch of the rank")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
hs, ws = synth_sizes(2, args.vignettes, 64, 1024)
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
n_batches = (args.vignettes + args.batch - 1) // args.batch
mine = list(range(rank, n_batches, world))
# size every lane's workspace for the largest batch of the job up front: a lane that has to grow its buffers inside a
# step pays a cudaMalloc / cudaFree pair of gigabytes (the step then takes milliseconds instead of one)
st.reserve([BatchGeometry(hs[b * args.batch:min((b + 1) * args.batch, args.vignettes)],
                          ws[b * args.batch:min((b + 1) * args.batch, args.vignettes)]) for b in mine])
ms = 0.0
n_vig = n_px = n_obj = checked = fallbacks = 0
for k, b in enumerate(mine):
    lo, hi = b * args.batch, min((b + 1) * args.batch, args.vignettes)
    g = BatchGeometry(hs[lo:hi], ws[lo:hi])
    db = st.prepare(DeviceBatch(g))
    img = db.synth(20261018, lo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = st.run_device(db, img)
    objs = r.n_obj
    st.join()
    e1.record()
    torch.cuda.synchronize()
    ms += e0.elapsed_time(e1)
    n_vig += g.n_img; n_px += g.pixels; n_obj += objs; fallbacks += len(r.dense_only)
    if k % args.check_every == 0:
        table = r.table
        used = g.total_px - 16          # (pad bytes between vignettes are zero-filled by the band kernel; 16 bytes of tail slack are not)
        fg = int(r.mask[:used].sum().item())
        assert int(table[:, 1].sum().item()) == fg, "areas do not add up to the mask"
        assert bool(((r.labels[:used] > 0) == (r.mask[:used] > 0)).all().item()), "labels != mask support"
        lab_off = r.lab_off.cpu().numpy()
        assert int(lab_off[-1]) == objs
        checked += 1
line = {"config": "configs[4]", "rank": rank, "world": world, "vignettes": n_vig, "mpix": n_px / 1e6, "objects": n_obj,
        "stage_ms": ms, "vignettes_per_s": n_vig / (ms / 1e3), "gpix_per_s": n_px / (ms / 1e3) / 1e9,
        "batches": len(mine), "batches_checked": checked, "per_operator_vignettes": fallbacks,
        "note": "one batch in flight at a time (generation and checks between steps): single-lane step time"}
print(json.dumps(line))
