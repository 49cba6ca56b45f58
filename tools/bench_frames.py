"""Timings for BASELINE.json configs[2] and configs[3] (single large frames, per-operator kernels), resident
on the device, CUDA events, best of 5.  Prints one JSON line per measurement (kept under profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maze_image_processing_pipeline_b200.device import WIDE_MIN_R, BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_dense_frame


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


# configs[2]: isotropic closing / opening, r = 1..32 on a 2048 x 2048 frame
frame = synth_dense_frame(7, size=2048, n_blobs=60)
g = BatchGeometry([2048], [2048]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
bits, flags = b.threshold_pack(d, 40)
for r in (1, 2, 4, 8, 16, 32, 40, 64):
    for op in ("closing", "opening"):
        ms, _ = timed(lambda: getattr(b, op)(bits, flags, r))
        print(json.dumps({"config": "configs[2]", "op": f"isotropic_{op}", "radius": r, "frame": "2048x2048",
                          "ms": round(ms, 4), "mpix_per_s": round(2048 * 2048 / ms / 1e3, 1),
                          "path": "bit-plane disk" if r < WIDE_MIN_R else "separable vertical distance + row test"}))
# configs[3]: 4096 x 4096 dense frame: threshold -> label -> regionprops (thousands of labels)
frame = synth_dense_frame(11, size=4096, n_blobs=3000)
g = BatchGeometry([4096], [4096]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
ms_t, (bits, flags) = timed(lambda: b.threshold_pack(d, 40))
ms_l, (labels, lab_off) = timed(lambda: b.label(bits))
n = int(lab_off[-1].item())
ms_p, table = timed(lambda: b.regionprops(lab_off, n, labels=labels, bits=bits, image=d, runs=True))
ms_u, _ = timed(lambda: b.unpack_mask(bits))
px = 4096 * 4096
print(json.dumps({"config": "configs[3]", "frame": "4096x4096", "labels": n, "ms_threshold_pack": round(ms_t, 4),
                  "ms_label": round(ms_l, 4), "ms_regionprops": round(ms_p, 4), "ms_unpack_mask": round(ms_u, 4),
                  "ms_total": round(ms_t + ms_l + ms_p + ms_u, 4),
                  "mpix_per_s": round(px / (ms_t + ms_l + ms_p + ms_u) / 1e3, 1),
                  "label_GBps_5B_per_px": round(5 * px / ms_l / 1e6, 1)}))

# configs[3] through the stage (band pipeline: band front, global-memory labelling of the frame, dense outputs)
from maze_image_processing_pipeline_b200 import stage as S
for name, pp in (("threshold+label+regionprops", S.SegmentationPostprocessingConfig()),
                 ("threshold+opening1+closing2+label+regionprops", S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1))):
    for compact in (False, True):
        st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=compact)
        db = st.prepare(DeviceBatch(g))
        st.reserve([g])
        for _ in range(st.n_lanes + 1):
            st.run_device(db, d).n_obj
        def run():
            rs = [st.run_device(db, d) for _ in range(4)]
            for r in rs:
                r.n_obj
            st.join()
            return rs[-1]
        ms, r = timed(run)
        ms /= 4
        print(json.dumps({"config": "configs[3]", "frame": "4096x4096", "path": "stage, band pipeline" + (" (compact)" if compact else ""),
                          "chain": name, "labels": int(r.n_obj), "redone": bool(r.redone), "ms_per_frame": round(ms, 4),
                          "mpix_per_s": round(px / ms / 1e3, 1), "stage_GBps_6B_per_px": round(6 * px / ms / 1e6, 1)}))
