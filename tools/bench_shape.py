"""Timing of maze_label_shape (perimeter / filled_area / euler_number per object) on a configs[1] batch of 2048
vignettes and on the configs[3] 4096^2 frame, resident on the device, CUDA events, best of 5.  One JSON line each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_dense_frame


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


hs, ws = bench.job_sizes()
B = 2048
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
g = BatchGeometry(hs[:B], ws[:B])
db = st.prepare(DeviceBatch(g))
img = db.synth(1, 0)
res = st.run_device(db, img)
table = res.table
torch.cuda.synchronize()
ms_stage, _ = timed(lambda: (st.run_device(db, img).table, st.join()))
ms_px, shape_px = timed(lambda: db.label_shape(table, labels=res.labels))
ms, shape = timed(lambda: db.label_shape(table, labels=res.labels, bits=res.bits, runs=True))
assert torch.equal(torch.nan_to_num(shape), torch.nan_to_num(shape_px))
sh = shape.cpu().numpy()
ok = ~np.isnan(sh[:, 0])
print(json.dumps({"config": "configs[1] batch", "vignettes": B, "mpix": round(g.total_px / 1e6, 1), "objects": int(ok.sum()),
                  "ms_label_shape": round(ms, 4), "ms_label_shape_per_pixel_planes": round(ms_px, 4),
                  "objects_with_holes": int((sh[ok, 1] > table.cpu().numpy()[ok, 1]).sum()),
                  "mean_perimeter": round(float(sh[ok, 0].mean()), 2)}))

frame = synth_dense_frame(11, size=4096, n_blobs=3000)
g = BatchGeometry([4096], [4096]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
bits, flags = b.threshold_pack(d, 40)
labels, lab_off = b.label(bits)
n = int(lab_off[-1].item())
table = b.regionprops(lab_off, n, labels=labels, bits=bits, image=d, runs=True)
ms, shape = timed(lambda: b.label_shape(table, labels=labels, bits=bits, runs=True))
print(json.dumps({"config": "configs[3]", "frame": "4096x4096", "labels": n, "ms_label_shape": round(ms, 4)}))

# where does the time go: the largest bounding boxes vs the rest (configs[1] batch)
g = BatchGeometry(hs[:B], ws[:B])
db = st.prepare(DeviceBatch(g))
img = db.synth(1, 0)
res = st.run_device(db, img)
table = res.table.clone()
t = table.cpu().numpy()
area = (t[:, 4] - t[:, 2]) * (t[:, 5] - t[:, 3])
area[~(t[:, 1] > 0)] = 0
order = np.argsort(-area)
for name, keep in (("largest 16", order[:16]), ("largest 128", order[:128]), ("all but largest 128", order[128:]),
                   ("bbox <= 64x64", np.nonzero(area <= 4096)[0])):
    t2 = t.copy()
    mask = np.ones(len(t), bool); mask[keep] = False
    t2[mask, 1] = 0
    d2 = torch.from_numpy(t2).to(table.device)
    ms, _ = timed(lambda: db.label_shape(d2, labels=res.labels, bits=res.bits, runs=True))
    print(json.dumps({"subset": name, "objects": int((t2[:, 1] > 0).sum()), "ms_label_shape": round(ms, 4),
                      "max_bbox": [int(np.sqrt(area[keep].max()))] if len(keep) else None}))
