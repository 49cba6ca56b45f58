"""ctypes binding of ``libmaze_b200.so`` (the C-ABI declared in ``include/maze_b200.h``).

There is no CPU fallback: if the shared library is missing or a symbol cannot be resolved the
import of the product package's device modules fails with :class:`MazeLibraryError`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libmaze_b200.so")

MAZE_OK = 0
MAZE_ERR_CUDA = -1
MAZE_ERR_TYPEERROR = -2
MAZE_ERR_BADARG = -3
MAZE_ERR_CAPACITY = -4

TILE_WORDS = 256
MAX_DISK_RADIUS = 32
NFEAT = 64
NACC = 12
NEXT = 8
NSHAPE = 8  # MAZE_NSHAPE: perimeter, filled_area, euler_number, n1, n2, n3, convex_area, -
RP_HIGH_ORDER = 1
RP_RUNS = 2
FUSED_CAPS = (1024, 2560, 4096, 6144, 9216, 13312, 19456, 28320)  # MAZE_FUSED_CAPS of include/maze_b200.h
FUSED_NO_PROPS = 4
BAND_SINGLE_REGION = 8   # MAZE_BAND_SINGLE_REGION
BAND_PLANE_WORDS = 6144  # MAZE_BAND_PLANE_WORDS
HUGE_PX = 1 << 21        # vignettes from 2 MPix on are frames: labelled by the global-memory kernels
STEP_COMPACT = 1         # MAZE_STEP_COMPACT


class StepArgs(ctypes.Structure):
    """maze_step_args_t of include/maze_b200.h."""
    _fields_ = ([(k, ctypes.c_void_p) for k in (
        "vig", "img_list", "left_vig", "left_tiles", "left_idx", "left_tiles_full", "image", "intensity", "bits",
        "mask", "labels", "counts", "lab_off", "stage_counter", "acc_stage", "hi_stage", "ext_stage", "table",
        "scratch_plane", "scratch_flags", "scratch_parent", "scratch_tile_scan", "scratch_lab_off", "scratch_acc",
        "scratch_ext", "counts_host", "bands", "band_off", "runs", "run_pix", "run_stats", "band_out", "band_counters",
        "big_list", "huge_host", "gl_scratch", "band_done")]
        + [("class_off", ctypes.c_int32 * (len(FUSED_CAPS) + 1)), ("pass_t", ctypes.c_int32 * 4), ("pass_invert", ctypes.c_int32 * 4)]
        + [(k, ctypes.c_int32) for k in ("n_img", "left_n", "left_n_tiles", "left_n_tiles_full", "t_int", "n_pass",
                                         "flags", "stage_cap", "n_bands", "halo", "run_cap", "step_flags")]
        + [("total_px", ctypes.c_int64), ("huge_px", ctypes.c_int64), ("min_area", ctypes.c_int64), ("n_huge", ctypes.c_int32),
           ("clear_border", ctypes.c_int32)])


class MazeLibraryError(RuntimeError):
    pass


class MazeCudaError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".sh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "maze_b200.h"))
    if not force and os.path.exists(SO_PATH):
        so_m = os.path.getmtime(SO_PATH)
        if all(os.path.getmtime(s) <= so_m for s in srcs):
            return SO_PATH
    cmd = ["sh", os.path.join(CSRC, "build.sh")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return SO_PATH


_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_d = ctypes.c_double

# name -> argtypes (restype is int unless noted); order as in include/maze_b200.h
SIGNATURES = {
    "maze_threshold_pack": [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp],
    "maze_morph_pass": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp],
    "maze_morph_pass_wide": [_vp, _vp, _vp, _i, _vp, ctypes.c_longlong, _vp, ctypes.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "maze_unpack_mask": [_vp, _vp, _i, _vp, _i, _vp, _vp],
    "maze_footprint_register": [_i, _vp],
    "maze_edt_sq": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "maze_compare_pack": [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp],
    "maze_label": [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp],
    "maze_clear_border": [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp],
    "maze_remove_small_objects": [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i64, _vp],
    "maze_max_label": [_vp, _vp, _i, _vp, _i, _vp, _vp],
    "maze_regionprops": [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp],
    "maze_merge_labels": [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                          _vp],
    "maze_merge_labels_ex": [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _i, _i, _vp],
    "maze_synth_vignettes": [_vp, _vp, _i, _vp, _i, _u64, _i64, _vp],
    "maze_vignette_stage": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i,
                            _vp, _vp, _vp, _vp],
    "maze_props_finish_staged": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "maze_count_scan": [_vp, _i, _vp, _vp],
    "maze_band_stage": [_vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                        _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, ctypes.c_longlong, _vp, _i, ctypes.c_longlong, _vp, _vp, _i, ctypes.c_longlong, _vp],
    "maze_label_shape": [_vp, _vp, _vp, _vp, _i, _vp, ctypes.c_longlong, _i, _i, _i, _vp, _vp, _vp],
    "maze_host_pack": [_vp, _vp, _vp, _i, _vp, _i],
    "maze_host_pack_wait": [_vp],
    "maze_host_expand": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i],
    "maze_host_expand_crop": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "maze_stage_step": [_vp, _vp, _vp],
    "maze_stage_step_graph": [_vp, _vp, _vp, _vp, _vp],
    "maze_graph_destroy": [_vp],
    "maze_front_chain": [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
}
OTHER_SYMBOLS = ["maze_host_pack_start", "maze_error_string", "maze_version", "maze_launch_count", "maze_prof_kernel_count",
                 "maze_prof_kernel_name", "maze_prof_enable", "maze_prof_collect"]

_lib = None


def lib():
    """The loaded library; raises MazeLibraryError when it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MazeLibraryError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the maze_b200 kernels)")
    try:
        handle = ctypes.CDLL(SO_PATH)
    except OSError as e:  # pragma: no cover
        raise MazeLibraryError(f"cannot load {SO_PATH}: {e}") from e
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise MazeLibraryError(f"{SO_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    handle.maze_host_pack_start.restype = ctypes.c_void_p
    handle.maze_host_pack_start.argtypes = [_vp, _vp, _vp, _i, _vp, _i]
    handle.maze_error_string.restype = ctypes.c_char_p
    handle.maze_version.restype = ctypes.c_int
    handle.maze_launch_count.restype = ctypes.c_longlong
    handle.maze_prof_kernel_name.restype = ctypes.c_char_p
    handle.maze_prof_kernel_name.argtypes = [ctypes.c_int]
    handle.maze_prof_enable.argtypes = [ctypes.c_int]
    handle.maze_prof_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc == MAZE_OK:
        return
    if rc == MAZE_ERR_CUDA:
        raise MazeCudaError(f"{what}: {lib().maze_error_string().decode()}")
    if rc == MAZE_ERR_BADARG:
        raise ValueError(f"{what}: bad argument")
    if rc == MAZE_ERR_CAPACITY:
        raise MemoryError(f"{what}: table capacity exceeded")
    raise RuntimeError(f"{what}: error {rc}")


def launch_count() -> int:
    return int(lib().maze_launch_count())


def prof_enable(on: bool):
    lib().maze_prof_enable(1 if on else 0)


def prof_collect():
    """{kernel name: (total ms, launches)} for everything recorded since the last collect."""
    import numpy as np
    n = lib().maze_prof_kernel_count()
    ms = np.zeros(n, np.float64)
    cnt = np.zeros(n, np.int64)
    check(lib().maze_prof_collect(ms.ctypes.data, cnt.ctypes.data, n), "maze_prof_collect")
    return {lib().maze_prof_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}
