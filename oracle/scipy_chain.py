"""The reference chain written against scipy.ndimage (TEST INFRASTRUCTURE ONLY).

The reference's arithmetic lives in scipy's C code (``distance_transform_edt``, ``label``,
``find_objects``), called from ``maze_ipp/isotropic.py`` and ``maze_ipp/merge_labels.py``.
``/root/reference`` does not exist on the GPU box, so this module restates those call
sequences against the same scipy entry points.  It therefore has the reference's CPU cost
profile and is what ``bench.py`` times as ``cpu_baseline`` (kind "port") and under
``--impl reference``.  It is checked against the real reference files in
``tests/golden/make_golden.py`` and against the independent C restatement in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage as ndi

from . import regionprops_table as _c_regionprops

_EIGHT = np.ones((3, 3), dtype=bool)


def _edt_cmp(fg, radius, keep_far, out=None):
    # isotropic.py:35-36 (keep_far: dist > radius) and :66-67 (dist < radius, strict)
    dist = ndi.distance_transform_edt(fg)
    return np.greater(dist, radius, out=out) if keep_far else np.less(dist, radius, out=out)


def erosion(mask, radius, out=None):
    return _edt_cmp(mask, radius, True, out)


def dilation(mask, radius, out=None):
    return _edt_cmp(mask == 0, radius, False, out)


def opening(mask, radius, out=None):  # isotropic.py:97-98
    return dilation(erosion(mask, radius, out=out), radius, out=out)


def closing(mask, radius, out=None):  # isotropic.py:128-129
    return erosion(dilation(mask, radius, out=out), radius, out=out)


def _apply_footprint(op, img, footprint, **kw):
    # skimage applies a footprint sequence element by element, each n times
    if isinstance(footprint, np.ndarray):
        return op(img, structure=footprint, **kw)
    for fp, n in footprint:
        for _ in range(int(n)):
            img = op(img, structure=fp, **kw)
    return img


def binary_erosion(mask, footprint):
    """skimage.morphology.binary_erosion: ndi.binary_erosion(..., border_value=True)."""
    return _apply_footprint(ndi.binary_erosion, np.asarray(mask) != 0, footprint, border_value=True)


def binary_dilation(mask, footprint):
    """skimage.morphology.binary_dilation: ndi.binary_dilation (outside = background)."""
    return _apply_footprint(ndi.binary_dilation, np.asarray(mask) != 0, footprint)


def binary_opening(mask, footprint):  # loki/pipeline.py:408-416
    return binary_dilation(binary_erosion(mask, footprint), footprint)


def binary_closing(mask, footprint):  # loki/pipeline.py:419-427
    return binary_erosion(binary_dilation(mask, footprint), footprint)


def label(mask):
    """loki/pipeline.py:430-433: skimage.measure.label(bool) == ndi.label with the full 3x3."""
    lab, n = ndi.label(mask, structure=_EIGHT)
    return lab.astype(np.int32, copy=False), int(n)


def clear_border(labels):  # loki/pipeline.py:435-439, in place
    edge = np.concatenate([labels[0, :], labels[-1, :], labels[:, 0], labels[:, -1]])
    hit = np.unique(edge[edge > 0])
    if hit.size:
        labels[np.isin(labels, hit)] = 0
    return labels


def remove_small_objects(labels, min_size):  # loki/pipeline.py:442-448, in place
    sizes = np.bincount(labels.ravel())
    small = sizes < min_size
    small[0] = False
    labels[small[labels]] = 0
    return labels


def _window_dist(mask, pad):
    # merge_labels.py:12-26
    if pad is None:
        return ndi.distance_transform_edt(~mask)
    (box,) = ndi.find_objects(mask, 1)
    box = tuple(slice(max(0, s.start - (pad + 1)), s.stop + pad + 1) for s in box)  # raises TypeError on None
    inner = ndi.distance_transform_edt(~mask[box])
    full = np.full(mask.shape, inner.max())
    full[box] = inner
    return full


def merge_labels(labels, index=None, max_distance=None, path_tolerance=5, return_merge_distances=False, labels_out=None):
    """merge_labels.py:29-113 restated: one seed cluster, nearest-first, stop at the first far label."""
    if index is None:
        u = np.unique(labels)
        index = u[u > 0].tolist()
    if len(index) < 2:
        return (labels, []) if return_merge_distances else labels
    if labels_out is None:
        labels_out = labels.copy()
    seed = index.pop(0)
    seed_mask = labels == seed
    labels_out[seed_mask] = seed
    pad = math.ceil(max_distance) if max_distance is not None else None
    near = _window_dist(seed_mask, pad)
    far = near.max()
    dists = []
    while index:
        k = int(np.argmin([near[labels == l].min(initial=far) for l in index]))
        cur = index.pop(k)
        cur_mask = labels == cur
        cur_near = _window_dist(cur_mask, pad)
        both = near + cur_near
        d = both.min()
        if max_distance is not None and d > max_distance:
            break
        bridge = cur_mask | (both <= d + path_tolerance)
        dists.append(d)
        labels_out[bridge] = seed
        closer = cur_near < near
        near[closer] = cur_near[closer]
    return (labels_out, dists) if return_merge_distances else labels_out


_REF = None


def reference_modules():
    """(isotropic, merge_labels) modules of the REAL reference from oracle/_ref/ (put there by oracle/make_ref.sh,
    see its header), or None when they are not present."""
    global _REF
    if _REF is None:
        import importlib.util
        import os
        mods = []
        for name in ("isotropic", "merge_labels"):
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", name + ".py")
            if not os.path.exists(path):
                mods = False
                break
            spec = importlib.util.spec_from_file_location("maze_ref_" + name, path)
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods.append(m)
        _REF = tuple(mods) if mods else False
    return _REF or None


def loki_chain(image, threshold_brighter=40, opening_radius=1, closing_radius=2, clear_border_flag=False,
               min_area=0, merge_segments_distance=0, with_props=True, use_reference=False):
    """One vignette through threshold -> opening -> closing -> label -> filters -> merge -> regionprops
    in the order of loki/pipeline.py:405-457 (opening BEFORE closing).  use_reference: morphology and merge_labels
    are the reference's own functions (oracle/_ref/), not the restatements above."""
    ref = reference_modules() if use_reference else None
    if use_reference and ref is None:
        raise RuntimeError("oracle/_ref/ is empty: run oracle/make_ref.sh where /root/reference exists")
    _opening = ref[0].isotropic_opening if ref else opening
    _closing = ref[0].isotropic_closing if ref else closing
    _merge = ref[1].merge_labels if ref else merge_labels
    mask = image > threshold_brighter
    if opening_radius > 0:
        mask = _opening(mask, opening_radius)
    if closing_radius > 0:
        mask = _closing(mask, closing_radius)
    labels, _ = label(mask)
    if clear_border_flag:
        clear_border(labels)
    if min_area > 0:
        remove_small_objects(labels, min_area)
    if merge_segments_distance > 0:
        labels = _merge(labels, max_distance=merge_segments_distance, labels_out=labels)
    table = _c_regionprops(np.ascontiguousarray(labels), image) if with_props else None
    return mask, labels, table
