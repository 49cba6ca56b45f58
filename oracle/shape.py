"""CPU restatement of the shape features skimage's RegionProperties computes from a region's crop
(TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and nothing else).

These are the values morphocut's ``CalculateZooProcessFeatures(region, meta, prefix="object_")`` reads at
maze_ipp/loki/pipeline.py:625 and :654 besides the moments.

PARITY UNPINNED for the skimage formulas: scikit-image is not installed here and is not vendored under
/root/reference (environment.yaml:11 lists it without a version), so ``perimeter``, ``euler_number`` and
``convex_area`` below restate the published algorithms of ``skimage.measure.perimeter`` /
``skimage.measure.euler_number`` / ``skimage.morphology.convex_hull_image`` (0.19-0.25) around the SAME scipy
primitives skimage calls (``ndimage.binary_erosion``, ``convolve``, ``binary_fill_holes``, ``find_objects``,
``spatial.ConvexHull``), which are executed, not restated.  tests/test_oracle.py additionally pins them to independent
definitions: the Euler number to (#8-connected components - #4-connected holes) counted with ``ndi.label``, the
perimeter to closed forms for rectangles, lines and single pixels.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage as ndi

_STREL_4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
_PERIMETER_KERNEL = np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]])
_PERIMETER_WEIGHTS = np.zeros(50, dtype=np.float64)
_PERIMETER_WEIGHTS[[5, 7, 15, 17, 25, 27]] = 1
_PERIMETER_WEIGHTS[[21, 33]] = math.sqrt(2)
_PERIMETER_WEIGHTS[[13, 23]] = (1 + math.sqrt(2)) / 2
_EULER_CONFIG = np.array([[0, 0, 0], [0, 1, 4], [0, 2, 8]])
_EULER_COEFS_8 = np.array([0, 0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0, -1, 0])


def perimeter_histogram(image) -> np.ndarray:
    """Histogram of the neighbourhood codes of skimage.measure.perimeter(image, neighborhood=4)."""
    image = np.asarray(image).astype(np.uint8)
    eroded = ndi.binary_erosion(image, _STREL_4, border_value=0)
    border = image - eroded
    codes = ndi.convolve(border, _PERIMETER_KERNEL, mode="constant", cval=0)
    return np.bincount(codes.ravel(), minlength=50)


def perimeter(image) -> float:
    """skimage.measure.perimeter(image, neighborhood=4) -- what RegionProperties.perimeter returns for
    ``region.image``."""
    return float(perimeter_histogram(image) @ _PERIMETER_WEIGHTS)


def perimeter_classes(image):
    """(n1, n2, n3): border pixels of weight 1, sqrt(2), (1 + sqrt(2)) / 2."""
    h = perimeter_histogram(image)
    return int(h[[5, 7, 15, 17, 25, 27]].sum()), int(h[[21, 33]].sum()), int(h[[13, 23]].sum())


def euler_number(image) -> int:
    """skimage.measure.euler_number(image, connectivity=2) for a 2-D image -- RegionProperties.euler_number."""
    image = np.pad(np.asarray(image) != 0, 1, mode="constant").astype(int)
    codes = ndi.convolve(image, _EULER_CONFIG, mode="constant", cval=0)
    h = np.bincount(codes.ravel(), minlength=16)
    return int(_EULER_COEFS_8 @ h)


def filled_area(image) -> int:
    """RegionProperties.filled_area: np.sum(ndi.binary_fill_holes(region.image, np.ones((3, 3))))."""
    return int(ndi.binary_fill_holes(np.asarray(image) != 0, np.ones((3, 3))).sum())


def label_shape(labels, max_label=None) -> np.ndarray:
    """One row per label 1..max_label: perimeter, filled_area, euler_number, n1, n2, n3, convex_area (NaN for absent labels),
    each taken from the label's own bounding-box crop ``labels[slice] == label`` as RegionProperties does."""
    labels = np.asarray(labels)
    n = int(labels.max()) if max_label is None else int(max_label)
    out = np.full((n, 8), np.nan)
    for lab, sl in enumerate(ndi.find_objects(labels, max_label=n), start=1):
        if sl is None:
            continue
        img = labels[sl] == lab
        n1, n2, n3 = perimeter_classes(img)
        out[lab - 1, :7] = (perimeter(img), filled_area(img), euler_number(img), n1, n2, n3, convex_area(img))
    return out


def convex_area(image) -> int:
    """RegionProperties.convex_area: np.sum(skimage.morphology.convex_hull_image(region.image)) as published for
    scikit-image >= 0.19 (offset_coordinates=True, include_borders=True): every object pixel contributes the four
    midpoints of its edges, scipy.spatial.ConvexHull (Qhull, executed) gives the hull, and a pixel belongs to the hull
    image when its centre lies inside the polygon or on its border.  Vertex coordinates are multiples of 1/2, so the
    cross products below are exact in float64."""
    from scipy.spatial import ConvexHull
    img = np.asarray(image) != 0
    if not img.any():
        return 0
    coords = np.argwhere(img).astype(np.float64)
    offsets = np.array([[0.5, 0.0], [-0.5, 0.0], [0.0, 0.5], [0.0, -0.5]])
    pts = np.unique((coords[:, None, :] + offsets[None, :, :]).reshape(-1, 2), axis=0)
    hull = ConvexHull(pts)
    v = pts[hull.vertices]                       # counter-clockwise in 2-D
    nxt = np.roll(v, -1, axis=0)
    e = (nxt - v)[None, :, :]
    total = 0
    step = max(1, (1 << 22) // max(1, img.shape[1] * len(v)))   # rows per chunk: bounded temporaries
    for y0 in range(0, img.shape[0], step):
        yy, xx = np.mgrid[y0:min(y0 + step, img.shape[0]), 0:img.shape[1]]
        p = np.stack([yy.ravel(), xx.ravel()], axis=1).astype(np.float64)
        d = p[:, None, :] - v[None, :, :]
        cross = e[:, :, 0] * d[:, :, 1] - e[:, :, 1] * d[:, :, 0]
        total += int(((cross >= 0).all(axis=1) | (cross <= 0).all(axis=1)).sum())
    return total
