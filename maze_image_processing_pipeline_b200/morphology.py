"""Drop-ins for the ``skimage.morphology`` callables of the LIVE post-processing chain
(maze_ipp/loki/pipeline.py:408-427): ``binary_opening(mask, disk(r, decomposition="crosses"))`` and
``binary_closing(...)``, plus ``binary_erosion`` / ``binary_dilation`` and ``disk`` themselves.

A footprint may be an ndarray or, as skimage allows, a sequence of ``(ndarray, n_iterations)`` pairs.  Erosion
(dilation) by a sequence equals erosion (dilation) by the Minkowski sum of its elements, so every footprint is
first collapsed into ONE array; the GPU kernels take it as a table of half chords per row (``maze_morph_pass``
with ``MAZE_FOOTPRINT_T``), which covers the symmetric, row-convex footprints with chords that do not grow away
from the centre row -- disks, squares, diamonds, crosses and everything ``disk(r, "crosses")`` expands to -- up
to a radius of 32.  Border rules are skimage's: pixels outside the image count as foreground for the erosion
(``border_value=True``) and as background for the dilation.

``disk(r, decomposition="crosses")`` restates ``skimage.morphology.footprints._cross_decomposition`` (scikit-image is
not available offline; SURVEY.md section 0.3 holds the same restatement and its check: the expansion is the closed
disk for r <= 12, 14-16, 19, 20 and differs from it by 8-16 pixels for r = 13, 17, 18).  Parity of the MORPHOLOGY is
pinned to ``scipy.ndimage.binary_erosion`` / ``binary_dilation`` with explicit structure arrays (tests).
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch

from ._lib import MAX_DISK_RADIUS, check, lib
from .device import BatchGeometry, DeviceBatch

Footprint = Union[np.ndarray, Sequence[Tuple[np.ndarray, int]]]


# ---- footprints (host side, tiny) ---------------------------------------------------------------------------
def _cross(r0: int, r1: int, dtype=np.uint8) -> np.ndarray:
    c = np.zeros((2 * r0 + 1, 2 * r1 + 1), dtype=dtype)
    if r0 == r1 == 0:
        c[:] = 1
        return c
    c[r0, :] = 1
    c[:, r1] = 1
    return c


def _cross_decomposition(footprint: np.ndarray, dtype=np.uint8):
    """Sequence of cross-shaped elements whose Minkowski sum approximates a symmetric convex footprint (after
    Li & Ritter; skimage.morphology.footprints._cross_decomposition, restated)."""
    quadrant = footprint[footprint.shape[0] // 2:, footprint.shape[1] // 2:]
    col_sums = np.concatenate((quadrant.sum(0, dtype=int), np.asarray([0], dtype=int)))
    i_prev, idx, sum0 = 0, {}, 0
    for i in range(col_sums.size - 1):
        if col_sums[i] > col_sums[i + 1]:
            if i == 0:
                continue
            key = (int(col_sums[i_prev] - col_sums[i]), i - i_prev)
            sum0 += key[0]
            idx[key] = idx.get(key, 0) + 1
            i_prev = i
    n = quadrant.shape[0] - 1 - sum0
    if n > 0:
        idx[(n, 0)] = idx.get((n, 0), 0) + 1
    return tuple((_cross(r0, r1, dtype), n) for (r0, r1), n in idx.items())


def disk(radius: int, dtype=np.uint8, *, strict_radius: bool = True, decomposition=None):
    """``skimage.morphology.disk``: pixels with x^2 + y^2 <= radius^2 (``strict_radius=False``: (radius + 0.5)^2);
    ``decomposition="crosses"`` returns the cross sequence the LOKI pipeline passes (pipeline.py:411-414)."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    rr = radius + (0 if strict_radius else 0.5)
    fp = np.array((X ** 2 + Y ** 2) <= rr ** 2, dtype=dtype)
    if decomposition is None:
        return fp
    if decomposition == "crosses":
        return _cross_decomposition(fp, dtype)
    raise NotImplementedError(f"decomposition={decomposition!r}")


def collapse(footprint: Footprint) -> np.ndarray:
    """One boolean array for a footprint or a footprint sequence (Minkowski sum of the elements)."""
    if isinstance(footprint, np.ndarray):
        return footprint != 0
    out = np.ones((1, 1), bool)
    for fp, n in footprint:
        fp = np.asarray(fp) != 0
        if fp.shape[0] % 2 == 0 or fp.shape[1] % 2 == 0:
            raise NotImplementedError("footprints with an even side (off-centre origin)")
        for _ in range(int(n)):
            ry, rx = fp.shape[0] // 2, fp.shape[1] // 2
            big = np.zeros((out.shape[0] + 2 * ry, out.shape[1] + 2 * rx), bool)
            for dy, dx in zip(*np.nonzero(fp)):
                big[dy:dy + out.shape[0], dx:dx + out.shape[1]] |= out
            out = big
    return out


def chord_table(footprint: Footprint) -> np.ndarray:
    """Half chords w[|dy|], |dy| = 0..R, of a footprint the kernels support (see the module docstring)."""
    fp = collapse(footprint)
    if fp.ndim != 2 or fp.shape[0] % 2 == 0 or fp.shape[1] % 2 == 0:
        raise NotImplementedError("footprints need odd sides (centred origin)")
    rows = np.nonzero(fp.any(1))[0]
    if rows.size == 0:
        raise ValueError("empty footprint")
    cy, cx = fp.shape[0] // 2, fp.shape[1] // 2
    R = int(max(cy - rows[0], rows[-1] - cy))
    w = np.zeros(R + 1, np.int32)
    for dy in range(-R, R + 1):
        row = fp[cy + dy] if 0 <= cy + dy < fp.shape[0] else np.zeros(fp.shape[1], bool)
        xs = np.nonzero(row)[0]
        if xs.size == 0 or xs[0] + xs[-1] != 2 * cx or xs.size != xs[-1] - xs[0] + 1:
            raise NotImplementedError("footprint rows must be centred runs without gaps")
        half = int(xs[-1] - cx)
        if dy < 0:
            w[-dy] = half
        elif w[dy] != half and dy > 0:
            raise NotImplementedError("footprint must be symmetric about its centre row")
        else:
            w[dy] = half
    if (np.diff(w) > 0).any() or R > MAX_DISK_RADIUS or w.max() > MAX_DISK_RADIUS:
        raise NotImplementedError(f"chords must not grow away from the centre row; radius <= {MAX_DISK_RADIUS}")
    return w


_PASS_CODES = {}
_PASS_RADIUS = {}  # pass code -> vertical radius R of the footprint


def pass_radius(t: int) -> int:
    """Rows a pass reaches up and down: isqrt(t) for a squared-distance threshold, R for a registered footprint."""
    import math
    t = int(t)
    if t >= 0:
        return math.isqrt(t)
    if t == -1:
        return 0
    return _PASS_RADIUS[t]


def footprint_pass_code(footprint: Footprint) -> int:
    """The pass threshold code (MAZE_FOOTPRINT_T(id)) under which the library knows this footprint."""
    w = chord_table(footprint)
    key = w.tobytes()
    if key not in _PASS_CODES:
        fid = lib().maze_footprint_register(len(w) - 1, w.ctypes.data)
        if fid < 0:
            check(fid, "maze_footprint_register")
        _PASS_CODES[key] = -2 - fid
        _PASS_RADIUS[-2 - fid] = len(w) - 1
    return _PASS_CODES[key]


# ---- per-image callables ------------------------------------------------------------------------------------
def _run(image, footprint, out, inverts):
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("2-D images only")
    if footprint is None:
        footprint = _cross(1, 1)  # skimage's default: the 4-neighbourhood cross
    code = footprint_pass_code(footprint)
    res_shape = img.shape
    if img.size == 0:
        res = np.zeros(res_shape, bool)
    else:
        fg = np.ascontiguousarray(img != 0).view(np.uint8)
        geom = BatchGeometry([fg.shape[0]], [fg.shape[1]])
        batch = DeviceBatch(geom)
        bits, flags = batch.threshold_pack(batch.upload(geom.pack_host([fg])), 0)
        for inv in inverts:
            bits, flags = batch.morph_pass(bits, flags, code, inv)
        res = geom.view(batch.unpack_mask(bits).cpu().numpy(), 0).astype(bool)
    if out is not None:
        out[...] = res
        return out
    return res


def binary_erosion(image, footprint: Footprint = None, out=None):
    """``skimage.morphology.binary_erosion`` (``ndi.binary_erosion(..., border_value=True)``)."""
    return _run(image, footprint, out, (0,))


def binary_dilation(image, footprint: Footprint = None, out=None):
    """``skimage.morphology.binary_dilation`` (outside the image is background)."""
    return _run(image, footprint, out, (1,))


def binary_opening(image, footprint: Footprint = None, out=None):
    """loki/pipeline.py:408-416: erosion followed by dilation with the same footprint."""
    return _run(image, footprint, out, (0, 1))


def binary_closing(image, footprint: Footprint = None, out=None):
    """loki/pipeline.py:419-427: dilation followed by erosion with the same footprint."""
    return _run(image, footprint, out, (1, 0))
