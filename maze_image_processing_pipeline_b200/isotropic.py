"""Drop-in for ``maze_ipp/isotropic.py`` (same names, arguments and return conventions), computed on
the GPU through the C-ABI (``maze_threshold_pack`` -> ``maze_morph_pass`` / ``maze_edt_sq`` ->
``maze_unpack_mask``).  Bit-exact with the reference, including its strict ``<`` in the dilation
(maze_ipp/isotropic.py:67) and scipy's phantom background pixel for planes without background.

These per-image callables pay one host<->device round trip each; the batched stage
(:mod:`.stage`) keeps whole batches resident instead.
"""
from __future__ import annotations

import numpy as np
import torch

from .device import BatchGeometry, DeviceBatch


def _as_plane(image):
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("isotropic morphology is implemented for 2-D images")
    fg = np.ascontiguousarray(img != 0).view(np.uint8)
    if fg.size == 0:
        return None, fg
    geom = BatchGeometry([fg.shape[0]], [fg.shape[1]])
    batch = DeviceBatch(geom)
    d_img = batch.upload(geom.pack_host([fg]))
    bits, flags = batch.threshold_pack(d_img, 0)
    return batch, (bits, flags)


def _finish(batch, bits, shape, out):
    mask = batch.unpack_mask(bits)
    res = batch.g.view(mask.cpu().numpy(), 0).astype(bool)
    if out is not None:
        out[...] = res
        return out
    return res


def _run(image, radius, out, ops):
    batch, state = _as_plane(image)
    if batch is None:
        res = np.zeros(np.asarray(image).shape, bool)
        if out is not None:
            out[...] = res
            return out
        return res
    bits, flags = state
    for op in ops:
        bits, flags = getattr(batch, op)(bits, flags, radius)
    return _finish(batch, bits, np.asarray(image).shape, out)


def isotropic_erosion(image, radius, out=None):
    """maze_ipp/isotropic.py:8-36: ``distance_transform_edt(image) > radius``."""
    return _run(image, radius, out, ("erosion",))


def isotropic_dilation(image, radius, out=None):
    """maze_ipp/isotropic.py:39-67: ``distance_transform_edt(image == 0) < radius``."""
    return _run(image, radius, out, ("dilation",))


def isotropic_opening(image, radius, out=None):
    """maze_ipp/isotropic.py:70-98: erosion followed by dilation."""
    return _run(image, radius, out, ("erosion", "dilation"))


def isotropic_closing(image, radius, out=None):
    """maze_ipp/isotropic.py:101-129: dilation followed by erosion."""
    return _run(image, radius, out, ("dilation", "erosion"))


def distance_transform_edt_sq(image):
    """Exact squared EDT (int32) with the semantics of ``scipy.ndimage.distance_transform_edt`` as
    used at maze_ipp/isotropic.py:35 (distance of nonzero pixels to the nearest zero pixel)."""
    batch, state = _as_plane(image)
    if batch is None:
        return np.zeros(np.asarray(image).shape, np.int32)
    d2 = batch.edt_sq(state[0], 0)
    return batch.g.view(d2.cpu().numpy(), 0).copy()
