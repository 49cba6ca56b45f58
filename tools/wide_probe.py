"""Per-kernel CUDA-event times of isotropic closing on a 2048 x 2048 frame for a few radii (maze_prof)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maze_image_processing_pipeline_b200 import _lib
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_dense_frame
frame = synth_dense_frame(7, size=2048, n_blobs=60)
g = BatchGeometry([2048], [2048]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
bits, flags = b.threshold_pack(d, 40)
for r in (4, 8, 32, 64):
    for _ in range(3):
        b.closing(bits, flags, r)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    t0 = time.perf_counter()
    for _ in range(10):
        b.closing(bits, flags, r)
    host = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    print(r, "host enqueue ms", round(host * 1e3, 4), {k: round(v[0] / 10, 4) for k, v in _lib.prof_collect().items()})
