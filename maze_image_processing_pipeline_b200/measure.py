"""GPU drop-ins for the skimage callables the LOKI post-processing chain invokes
(maze_ipp/loki/pipeline.py:430-448, 589-625, 653-654): ``label``, ``clear_border``,
``remove_small_objects`` and a ``regionprops_table`` carrying the RegionProperties subset the
stage reads.  Single-image host API; the batched stage keeps data on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import NFEAT, NSHAPE
from .device import BatchGeometry, DeviceBatch

# column layout of the feature table (include/maze_b200.h, MAZE_F_*)
F_LABEL, F_AREA, F_BBOX, F_CENTROID, F_MU, F_NU, F_HU, F_EIG = 0, 1, 2, 6, 8, 24, 40, 47
F_AXIS_MAJOR, F_AXIS_MINOR, F_ECC, F_ORIENT = 49, 50, 51, 52
F_IMIN, F_IMAX, F_IMEAN, F_FRAC_INVALID, F_IMAGE, F_T00, F_T01, F_T11 = 53, 54, 55, 56, 57, 58, 59, 60


# column layout of the shape table (MAZE_S_*)
S_PERIMETER, S_FILLED_AREA, S_EULER, S_N1, S_N2, S_N3, S_CONVEX_AREA = 0, 1, 2, 3, 4, 5, 6


def _single(arr, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    if a.ndim != 2:
        raise ValueError("2-D images only")
    geom = BatchGeometry([a.shape[0]], [a.shape[1]])
    batch = DeviceBatch(geom)
    return geom, batch, batch.upload(geom.pack_host([a], dtype=dtype))


def label(label_image, return_num=False):
    """loki/pipeline.py:430-433: ``skimage.measure.label`` of a bool mask (8-connectivity, labels in
    raster order of each component's first pixel, int32)."""
    m = np.asarray(label_image)
    if m.dtype != bool:
        raise TypeError("this drop-in labels boolean masks (the pipeline's call site)")
    if m.size == 0:
        res = np.zeros(m.shape, np.int32)
        return (res, 0) if return_num else res
    geom, batch, d_img = _single(m.view(np.uint8), np.uint8)
    bits, _ = batch.threshold_pack(d_img, 0)
    labels, lab_off = batch.label(bits)
    res = geom.view(labels.cpu().numpy(), 0).copy()
    if return_num:
        return res, int(lab_off.cpu()[1])
    return res


def _inplace_filter(labels, out, which, **kw):
    lab = np.asarray(labels)
    if lab.size == 0:
        return labels if out is None else out
    geom, batch, d_lab = _single(lab, np.int32)
    bound = int(batch.max_label(d_lab).cpu()[0])
    lab_off, n_obj = batch.lab_off_from_bounds([bound])
    getattr(batch, which)(d_lab, lab_off, n_obj, **kw)
    res = geom.view(d_lab.cpu().numpy(), 0)
    if out is None:
        return res.astype(lab.dtype, copy=True)
    out[...] = res
    return out


def clear_border(labels, out=None):
    """loki/pipeline.py:435-439 (buffer_size=0, bgval=0): labels touching the border become 0."""
    return _inplace_filter(labels, out, "clear_border")


def remove_small_objects(ar, min_size=64, out=None):
    """loki/pipeline.py:442-448 on a label image: labels with fewer than min_size pixels become 0."""
    return _inplace_filter(ar, out, "remove_small_objects", min_size=int(min_size))


def regionprops_table(labels, intensity_image=None, high_order=True):
    """(max_label, NFEAT) float64 table; row l-1 describes label l (area 0 / NaN when absent)."""
    lab = np.asarray(labels)
    if lab.size == 0:
        return np.zeros((0, NFEAT))
    geom, batch, d_lab = _single(lab, np.int32)
    d_img = None
    if intensity_image is not None:
        d_img = batch.upload(geom.pack_host([np.ascontiguousarray(intensity_image, dtype=np.uint8)]))
    bound = int(batch.max_label(d_lab).cpu()[0])
    lab_off, n_obj = batch.lab_off_from_bounds([bound])
    table = batch.regionprops(lab_off, n_obj, labels=d_lab, image=d_img, high_order=high_order)
    return table.cpu().numpy()


def mask_properties(mask, intensity_image=None, high_order=True):
    """loki/pipeline.py:653 ``ImageProperties(mask, image)``: the whole mask as one region."""
    m = np.asarray(mask)
    geom, batch, d_m = _single((m != 0).view(np.uint8), np.uint8)
    bits, _ = batch.threshold_pack(d_m, 0)
    d_img = None
    if intensity_image is not None:
        d_img = batch.upload(geom.pack_host([np.ascontiguousarray(intensity_image, dtype=np.uint8)]))
    lab_off, n_obj = batch.lab_off_from_bounds([1])
    return batch.regionprops(lab_off, n_obj, bits=bits, image=d_img, high_order=high_order).cpu().numpy()


def regionprops_shape(labels):
    """(max_label, NSHAPE) float64 table of the RegionProperties values that need the pixel neighbourhood --
    ``perimeter`` (4-neighbourhood), ``filled_area`` (holes filled with the full 3x3 structure),
    ``euler_number`` (8-connectivity) and ``convex_area`` (convex hull image), each taken from the label's own bounding-box crop -- as read by
    ``CalculateZooProcessFeatures`` (loki/pipeline.py:625).  Row l-1 describes label l (NaN when absent)."""
    lab = np.asarray(labels)
    if lab.size == 0:
        return np.zeros((0, NSHAPE))
    geom, batch, d_lab = _single(lab, np.int32)
    bound = int(batch.max_label(d_lab).cpu()[0])
    lab_off, n_obj = batch.lab_off_from_bounds([bound])
    table = batch.regionprops(lab_off, n_obj, labels=d_lab, high_order=False)
    return batch.label_shape(table, labels=d_lab).cpu().numpy()


def mask_shape(mask):
    """The same three values for the whole mask as one region (``ImageProperties``, loki/pipeline.py:653-654)."""
    m = np.asarray(mask)
    geom, batch, d_m = _single((m != 0).view(np.uint8), np.uint8)
    bits, _ = batch.threshold_pack(d_m, 0)
    lab_off, n_obj = batch.lab_off_from_bounds([1])
    table = batch.regionprops(lab_off, n_obj, bits=bits, high_order=False)
    return batch.label_shape(table, bits=bits).cpu().numpy()
