"""bench.py with another build of the library (experiments): python tools/lib_variant_bench.py libmaze_b200_t32.so [bench args]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maze_image_processing_pipeline_b200 import _lib
_lib.SO_PATH = os.path.join(_lib.CSRC, sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import bench
sys.exit(bench.main())
