"""B200-native LOKI re-segmentation stage of MAZE-IPP (threshold -> EDT-based isotropic opening /
closing, or the live pipeline's footprint morphology -> 8-connected labelling -> label filters / merge_labels ->
regionprops and ZooProcess shape features).

Host side: Python + PyTorch for device memory only; all computation is hand-written sm_100a CUDA in
``csrc/`` reached through the C-ABI of ``include/maze_b200.h`` (ctypes, :mod:`._lib`).  There is no
CPU fallback: using any operator without the built library or without a GPU raises.
"""
__version__ = "0.1.0"

__all__ = [
    "isotropic_erosion", "isotropic_dilation", "isotropic_opening", "isotropic_closing",
    "merge_labels", "label", "clear_border", "remove_small_objects", "regionprops_table", "regionprops_shape",
    "binary_erosion", "binary_dilation", "binary_opening", "binary_closing", "disk",
    "LokiSegmentationStage", "SegmentationPostprocessingConfig", "ThresholdSegmentationConfig",
]


def __getattr__(name):  # lazy: importing the package must work on a CPU-only box (tests of host logic)
    if name in ("isotropic_erosion", "isotropic_dilation", "isotropic_opening", "isotropic_closing"):
        from . import isotropic
        return getattr(isotropic, name)
    if name == "merge_labels":
        from .merge_labels import merge_labels
        return merge_labels
    if name in ("binary_erosion", "binary_dilation", "binary_opening", "binary_closing", "disk"):
        from . import morphology
        return getattr(morphology, name)
    if name in ("label", "clear_border", "remove_small_objects", "regionprops_table", "mask_properties",
                "regionprops_shape", "mask_shape"):
        from . import measure
        return getattr(measure, name)
    if name in ("LokiSegmentationStage", "SegmentationPostprocessingConfig", "ThresholdSegmentationConfig"):
        from . import stage
        return getattr(stage, name)
    raise AttributeError(name)
