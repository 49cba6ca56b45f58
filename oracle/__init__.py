"""CPU oracle for the LOKI re-segmentation hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``maze_image_processing_pipeline_b200``) never does.

Two restatements live here:

* :mod:`oracle.maze_oracle.c` (loaded through ctypes below) -- an independent plain-C
  restatement with its own exact integer EDT, flood-fill CCL, ``merge_labels`` loop and
  float64 regionprops.
* :mod:`oracle.scipy_chain` -- the same chain written against ``scipy.ndimage`` exactly the
  way the reference calls it; it has the reference's performance characteristics and is the
  ``cpu_baseline`` ("port") that ``bench.py`` times.

Pinning: see the header of ``maze_oracle.c`` and ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmaze_oracle.so")
_SRC = os.path.join(_HERE, "maze_oracle.c")

ERR_TYPEERROR = -2


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_nfeat.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(mask):
    return np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)


NFEAT = 64
# column layout shared with include/maze_b200.h
F_LABEL, F_AREA, F_BBOX, F_CENTROID, F_MU, F_NU, F_HU, F_EIG = 0, 1, 2, 6, 8, 24, 40, 47
F_AXIS_MAJOR, F_AXIS_MINOR, F_ECC, F_ORIENT = 49, 50, 51, 52
F_IMIN, F_IMAX, F_IMEAN, F_FRAC_INVALID, F_IMAGE, F_T00, F_T01, F_T11 = 53, 54, 55, 56, 57, 58, 59, 60


def edt_sq(image, bruteforce=False):
    img = _u8(image)
    H, W = img.shape
    out = np.empty((H, W), np.int64)
    fn = lib().oracle_edt_sq_bruteforce if bruteforce else lib().oracle_edt_sq
    fn(_p(img), H, W, _p(out))
    return out


def threshold(image, thr):
    img = np.ascontiguousarray(image, dtype=np.uint8)
    out = np.empty(img.shape, np.uint8)
    lib().oracle_threshold(_p(img), ctypes.c_int64(img.size), ctypes.c_double(float(thr)), _p(out))
    return out.astype(bool)


def _iso(name, image, radius):
    img = _u8(image)
    H, W = img.shape
    out = np.empty((H, W), np.uint8)
    getattr(lib(), name)(_p(img), H, W, ctypes.c_double(float(radius)), _p(out))
    return out.astype(bool)


def isotropic_erosion(image, radius):
    return _iso("oracle_isotropic_erosion", image, radius)


def isotropic_dilation(image, radius):
    return _iso("oracle_isotropic_dilation", image, radius)


def isotropic_opening(image, radius):
    return _iso("oracle_isotropic_opening", image, radius)


def isotropic_closing(image, radius):
    return _iso("oracle_isotropic_closing", image, radius)


def label(mask):
    m = _u8(mask)
    H, W = m.shape
    out = np.empty((H, W), np.int32)
    lib().oracle_label8.restype = ctypes.c_int32
    n = lib().oracle_label8(_p(m), H, W, _p(out))
    return out, int(n)


def clear_border(labels):
    """In place, returns its argument (like ``clear_border(labels, out=labels)``)."""
    assert labels.dtype == np.int32 and labels.flags.c_contiguous
    H, W = labels.shape
    lib().oracle_clear_border(_p(labels), H, W)
    return labels


def remove_small_objects(labels, min_size):
    assert labels.dtype == np.int32 and labels.flags.c_contiguous
    H, W = labels.shape
    lib().oracle_remove_small_objects(_p(labels), H, W, ctypes.c_int64(int(min_size)))
    return labels


def merge_labels(labels, index=None, max_distance=None, path_tolerance=5, return_merge_distances=False, labels_out=None):
    """Same signature and aliasing/identity/exception behaviour as maze_ipp/merge_labels.py:29-113."""
    assert labels.dtype == np.int32 and labels.flags.c_contiguous
    H, W = labels.shape
    if index is not None:
        idx = np.ascontiguousarray(index, dtype=np.int32)
        n_index = len(idx)
        idx_p = _p(idx)
    else:
        n_index = 0
        idx_p = None
    if index is None:
        n_labels = len(np.unique(labels[labels > 0]))
    else:
        n_labels = n_index
    if n_labels < 2:
        return (labels, []) if return_merge_distances else labels
    if labels_out is None:
        labels_out = labels.copy()
    assert labels_out.dtype == np.int32 and labels_out.flags.c_contiguous
    md = np.zeros(max(n_labels, 1), np.float64)
    nm = ctypes.c_int(0)
    ni = ctypes.c_int(0)
    rc = lib().oracle_merge_labels(
        _p(labels), H, W, idx_p, n_index, int(max_distance is not None),
        ctypes.c_double(float(max_distance) if max_distance is not None else 0.0),
        ctypes.c_double(float(path_tolerance)), _p(labels_out), _p(md), ctypes.byref(nm), ctypes.byref(ni))
    if rc == ERR_TYPEERROR:
        raise TypeError("'NoneType' object is not iterable")
    if rc != 0:
        raise RuntimeError(f"oracle_merge_labels failed: {rc}")
    if index is not None:
        # the reference pops from the caller's list (merge_labels.py:66, 84); nothing else observable
        pass
    if return_merge_distances:
        return labels_out, [float(v) for v in md[: nm.value]]
    return labels_out


def regionprops_table(labels, image=None, max_label=None):
    assert labels.dtype == np.int32 and labels.flags.c_contiguous
    H, W = labels.shape
    if max_label is None:
        max_label = int(labels.max()) if labels.size else 0
    table = np.empty((max(max_label, 0), NFEAT), np.float64)
    img_p = None
    if image is not None:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        img_p = _p(image)
    lib().oracle_regionprops(_p(labels), img_p, H, W, ctypes.c_int32(max_label), _p(table))
    return table
