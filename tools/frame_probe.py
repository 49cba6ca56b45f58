"""Per-kernel CUDA-event times of one 4096 x 4096 frame through the stage (band pipeline, global-memory labelling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maze_image_processing_pipeline_b200 import _lib, stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
from maze_image_processing_pipeline_b200.synth import synth_dense_frame
frame = synth_dense_frame(11, size=4096, n_blobs=3000)
g = BatchGeometry([4096], [4096]); b = DeviceBatch(g)
d = b.upload(g.pack_host([frame]))
for compact in (False, True):
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1),
                                 compact=compact, n_lanes=1)
    db = st.prepare(DeviceBatch(g)); st.reserve([g])
    for _ in range(3):
        r = st.run_device(db, d); r.n_obj
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(5):
        r = st.run_device(db, d); r.n_obj
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    print("compact" if compact else "dense", "runs", r.n_runs, "objects", r.n_obj,
          {k: (round(v[0] / 5, 4), v[1] // 5) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])})
