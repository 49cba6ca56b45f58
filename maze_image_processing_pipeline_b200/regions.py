"""Per-object metadata from the stage's object table: the host-side tail of the LOKI segmentation
(maze_ipp/loki/pipeline.py:589-625).

The reference runs, per frame, ``FindRegions(labels, image, padding=75, min_intensity=...)`` (one skimage
``RegionProperties`` per surviving label, its slice enlarged by ``padding`` with the start clipped at 0 and
the stop left unclipped), then ``recalc_metadata`` (:604-619) and ``CalculateZooProcessFeatures(region, meta,
prefix="object_")`` (:625).  Here the numbers come from the feature table the GPU produced
(``StageResult.features(i)``, columns ``MAZE_F_*`` of include/maze_b200.h); nothing is recomputed from pixels.

Restated, not executed: morphocut and scikit-image are not available offline, so the
key names follow morphocut's ZooProcess feature set as far as the stage's table carries the numbers; features
that need perimeter / convex hull / hole filling are not produced yet.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional

import numpy as np

from .measure import (F_AREA, F_AXIS_MAJOR, F_AXIS_MINOR, F_BBOX, F_CENTROID, F_ECC, F_FRAC_INVALID, F_HU, F_IMAX,
                      F_IMEAN, F_IMIN, F_LABEL, F_ORIENT)


class Region:
    """The part of skimage's RegionProperties the LOKI pipeline reads, backed by one row of the table."""

    def __init__(self, row: np.ndarray, shape, padding: int = 0, labels: Optional[np.ndarray] = None,
                 intensity: Optional[np.ndarray] = None):
        self._row = row
        self._shape = shape
        self._labels = labels
        self._intensity = intensity
        r0, c0, r1, c1 = (int(v) for v in row[F_BBOX:F_BBOX + 4])
        # _enlarge_slice semantics of morphocut's FindRegions: start clipped at 0, stop NOT clipped
        self.slice = (slice(max(0, r0 - padding), r1 + padding), slice(max(0, c0 - padding), c1 + padding))

    @property
    def label(self) -> int:
        return int(self._row[F_LABEL])

    @property
    def bbox(self):
        """(min_row, min_col, max_row, max_col) of the (padded) slice, as skimage reports it."""
        return (self.slice[0].start, self.slice[1].start, self.slice[0].stop, self.slice[1].stop)

    @property
    def area(self) -> float:
        return float(self._row[F_AREA])

    @property
    def centroid(self):
        return (float(self._row[F_CENTROID]), float(self._row[F_CENTROID + 1]))

    @property
    def intensity_max(self) -> float:
        return float(self._row[F_IMAX])

    @property
    def image(self) -> np.ndarray:
        """Boolean mask of the object inside the (padded, image-clipped) slice."""
        if self._labels is None:
            raise ValueError("label image not attached")
        return self._labels[self.slice] == self.label

    @property
    def image_intensity(self) -> np.ndarray:
        if self._intensity is None:
            raise ValueError("intensity image not attached")
        return self._intensity[self.slice] * self.image

    def __getitem__(self, col: int) -> float:
        return float(self._row[col])


def find_regions(result, i: int, padding: int = 0, min_intensity: Optional[float] = None,
                 image: Optional[np.ndarray] = None) -> Iterator[Region]:
    """FindRegions(labels, image, padding, min_intensity) for vignette / frame ``i`` of a StageResult
    (loki/pipeline.py:589-594): one Region per label that still has pixels, in label order; regions whose
    maximum intensity is below ``min_intensity`` are skipped."""
    feats = result.features(i)
    labels = result.labels(i)
    shape = (int(result.geometry.h[i]), int(result.geometry.w[i]))
    for row in feats:
        if not row[F_AREA] > 0:  # label removed by clear_border / remove_small_objects / merge_labels
            continue
        if min_intensity is not None and row[F_IMAX] < min_intensity:
            continue
        yield Region(row, shape, padding, labels, image)


def recalc_metadata(region: Region, meta: Dict, object_id_fmt: Optional[str] = None) -> Dict:
    """loki/pipeline.py:604-619, including its unpacking of ``region.bbox`` as ``(y0, x0, x1, y1)``: skimage's
    bbox is (min_row, min_col, max_row, max_col), so ``object_width`` is max_row - min_col and
    ``object_height`` is max_col - min_row, exactly as the reference computes them."""
    meta = dict(meta)
    (y0, x0, x1, y1) = region.bbox
    meta["object_posx"] = x0
    meta["object_posy"] = y0
    meta["object_sequence"] = region.label
    meta["object_width"] = x1 - x0
    meta["object_height"] = y1 - y0
    if object_id_fmt is not None:
        meta["object_id"] = object_id_fmt.format_map(meta)
    # (region.image_intensity[region.image] == 0).mean() -- taken from the exact zero count of the table
    meta["object_frac_invalid"] = float(region[F_FRAC_INVALID])
    return meta


def zooprocess_features(region: Region, meta: Optional[Dict] = None, prefix: str = "object_") -> Dict:
    """The subset of ``CalculateZooProcessFeatures(region, meta, prefix)`` (loki/pipeline.py:625, 654) that
    follows from the stage's table: area, intensity statistics, centroid, bounding box, ellipse axes, angle and
    ratios derived from them.  Keys follow morphocut's ZooProcess names."""
    out = dict(meta) if meta is not None else {}
    row = region._row
    r0, c0, r1, c1 = (int(v) for v in row[F_BBOX:F_BBOX + 4])
    area = float(row[F_AREA])
    major, minor = float(row[F_AXIS_MAJOR]), float(row[F_AXIS_MINOR])
    feats = {
        "area": area,
        "mean": float(row[F_IMEAN]),
        "min": float(row[F_IMIN]),
        "max": float(row[F_IMAX]),
        "x": float(row[F_CENTROID + 1]),
        "y": float(row[F_CENTROID]),
        "bx": c0,
        "by": r0,
        "width": c1 - c0,
        "height": r1 - r0,
        "major": major,
        "minor": minor,
        "angle": float(row[F_ORIENT]) / math.pi * 180.0 + 90.0,
        "intden": area * float(row[F_IMEAN]),
        "range": float(row[F_IMAX]) - float(row[F_IMIN]),
        "elongation": (major / minor) if minor > 0 else float("inf"),
        "eccentricity": float(row[F_ECC]),
    }
    for k, v in feats.items():
        out[prefix + k] = v
    for j in range(7):
        out[f"{prefix}hu{j + 1}"] = float(row[F_HU + j])
    return out


def objects_of(result, i: int, meta: Optional[Dict] = None, padding: int = 75, min_intensity: Optional[float] = None,
               image: Optional[np.ndarray] = None, object_id_fmt: Optional[str] = None) -> List[Dict]:
    """All objects of vignette / frame ``i`` as metadata dicts: FindRegions -> recalc_metadata ->
    CalculateZooProcessFeatures, in the order of loki/pipeline.py:589-625 (default padding 75,
    loki/config_schema.py:90-96)."""
    base = {} if meta is None else meta
    return [zooprocess_features(r, recalc_metadata(r, base, object_id_fmt))
            for r in find_regions(result, i, padding, min_intensity, image)]
