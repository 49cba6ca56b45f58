"""Two steps of the stage with merge_segments_distance = 10 on the bench batch (target for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
B = 4096
pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, merge_segments_distance=10)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, merge_errors="ignore")
g = BatchGeometry(hs[:B], ws[:B]); db = DeviceBatch(g)
img = db.synth(bench.PIXEL_SEED, 0)
for _ in range(2):
    print(st.run_device(db, img).n_obj)
torch.cuda.synchronize()
