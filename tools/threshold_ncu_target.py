"""Two steps of the threshold branch on the bench batch (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maze_image_processing_pipeline_b200 import stage as S
from maze_image_processing_pipeline_b200.device import BatchGeometry, DeviceBatch
hs, ws = bench.job_sizes()
g = BatchGeometry(hs[:4096], ws[:4096]); db = DeviceBatch(g); img = db.synth(bench.PIXEL_SEED, 0)
st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40))
for _ in range(2):
    print(st.run_device(db, img).n_obj)
torch.cuda.synchronize()
