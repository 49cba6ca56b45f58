"""Per-object metadata from the stage's object table: the host-side tail of the LOKI segmentation
(maze_ipp/loki/pipeline.py:589-625).

The reference runs, per frame, ``FindRegions(labels, image, padding=75, min_intensity=...)`` (one skimage
``RegionProperties`` per surviving label, its slice enlarged by ``padding`` with the start clipped at 0 and
the stop left unclipped), then ``recalc_metadata`` (:604-619) and ``CalculateZooProcessFeatures(region, meta,
prefix="object_")`` (:625).  Here the numbers come from the feature table the GPU produced
(``StageResult.features(i)``, columns ``MAZE_F_*`` of include/maze_b200.h); nothing is recomputed from pixels.

Restated, not executed: morphocut and scikit-image are not available offline, so the
key names follow morphocut's ZooProcess feature set (``morphocut/contrib/zooprocess.py`` at the pinned commit
03dbc6b, requirements.txt:1) as remembered: there ``area`` is the FILLED area and ``area_exc`` the pixel count.
With the stage's shape table (``LokiSegmentationStage(shape_features=True)``: perimeter, filled_area,
euler_number, convex_area from maze_label_shape) every key of that set is produced; without it, only the keys
that follow from the moment table.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional

import numpy as np

# column layouts (include/maze_b200.h, MAZE_F_* / MAZE_S_*); kept literal so that this module imports without CUDA
F_LABEL, F_AREA, F_BBOX, F_CENTROID, F_HU = 0, 1, 2, 6, 40
F_AXIS_MAJOR, F_AXIS_MINOR, F_ECC, F_ORIENT = 49, 50, 51, 52
F_IMIN, F_IMAX, F_IMEAN, F_FRAC_INVALID = 53, 54, 55, 56
S_PERIMETER, S_FILLED_AREA, S_EULER, S_CONVEX_AREA = 0, 1, 2, 6


class Region:
    """The part of skimage's RegionProperties the LOKI pipeline reads, backed by one row of the table."""

    def __init__(self, row: np.ndarray, shape, padding: int = 0, labels: Optional[np.ndarray] = None,
                 intensity: Optional[np.ndarray] = None, shape_row: Optional[np.ndarray] = None, mask_fn=None,
                 sequence: Optional[int] = None, whole_frame: bool = False, label_fn=None):
        self._row = row
        self._shape_row = shape_row
        self._shape = shape
        self._labels = labels
        self._intensity = intensity
        self._mask_fn = mask_fn  # (slice, label) -> bool crop; compact results expand the object from its runs
        self._label_fn = label_fn  # slice -> int32 crop of the label image (compact results: from the run list)
        self.sequence = int(row[F_LABEL]) if sequence is None else int(sequence)
        r0, c0, r1, c1 = (int(v) for v in row[F_BBOX:F_BBOX + 4])
        # _enlarge_slice semantics of morphocut's FindRegions: start clipped at 0, stop NOT clipped
        self.slice = (slice(max(0, r0 - padding), r1 + padding), slice(max(0, c0 - padding), c1 + padding))
        if whole_frame:  # ImageProperties(mask, image), loki/pipeline.py:653: the region's slice is the whole frame
            self.slice = (slice(0, int(shape[0])), slice(0, int(shape[1])))

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
    def perimeter(self) -> float:
        return float(self._need_shape()[S_PERIMETER])

    @property
    def filled_area(self) -> float:
        return float(self._need_shape()[S_FILLED_AREA])

    @property
    def euler_number(self) -> int:
        return int(self._need_shape()[S_EULER])

    @property
    def convex_area(self) -> float:
        return float(self._need_shape()[S_CONVEX_AREA])

    def _need_shape(self):
        if self._shape_row is None:
            raise ValueError("shape table not attached (LokiSegmentationStage(shape_features=True))")
        return self._shape_row

    @property
    def centroid(self):
        return (float(self._row[F_CENTROID]), float(self._row[F_CENTROID + 1]))

    @property
    def intensity_max(self) -> float:
        return float(self._row[F_IMAX])

    @property
    def crop_shape(self):
        """Shape of ``image``: the padded slice clipped to the frame (numpy clips the unclipped stop)."""
        h, w = self._shape
        return (min(self.slice[0].stop, h) - self.slice[0].start, min(self.slice[1].stop, w) - self.slice[1].start)

    @property
    def image(self) -> np.ndarray:
        """Boolean mask of the object inside the (padded, image-clipped) slice."""
        if self._labels is not None:
            return self._labels[self.slice] == self.label
        if self._mask_fn is not None:
            return self._mask_fn(self.slice, self.label)
        raise ValueError("label image not attached")

    @property
    def label_image(self) -> np.ndarray:
        """The label image inside the (padded, image-clipped) slice: what tells this object from the others."""
        if self._labels is not None:
            return self._labels[self.slice]
        if self._label_fn is not None:
            return self._label_fn(self.slice)
        raise ValueError("label image not attached")

    @property
    def image_intensity(self) -> np.ndarray:
        if self._intensity is None:
            raise ValueError("intensity image not attached")
        return self._intensity[self.slice] * self.image

    def __getitem__(self, col: int) -> float:
        return float(self._row[col])


def find_regions(result, i: int, padding: int = 0, min_intensity: Optional[float] = None,
                 image: Optional[np.ndarray] = None, renumber: bool = True) -> Iterator[Region]:
    """FindRegions(labels, image, padding, min_intensity) for vignette / frame ``i`` of a StageResult
    (loki/pipeline.py:589-594): one Region per label that still has pixels, in label order; regions whose
    maximum intensity is below ``min_intensity`` are skipped."""
    feats = result.features(i)
    shapes = result.shape_features(i) if hasattr(result, "shape_features") else None
    shape = (int(result.geometry.h[i]), int(result.geometry.w[i]))
    labels, mask_fn, label_fn, whole = None, None, None, False
    if getattr(result, "compact", False):  # run-list result: objects are expanded one crop at a time
        mask_fn = lambda sl, lab, _i=i: result.object_mask(_i, sl, lab)  # noqa: E731
        label_fn = lambda sl, _i=i: result.label_crop(_i, sl)  # noqa: E731
    else:
        labels = result.labels(i)
        if labels is None:  # threshold branch (ImageProperties): the whole mask is label 1
            mask_fn = lambda sl, lab, _i=i: result.mask(_i)[sl]  # noqa: E731
            whole = True
    seq = 0
    for j, row in enumerate(feats):
        if not row[F_AREA] > 0:  # label removed by clear_border / remove_small_objects / merge_labels
            continue
        # morphocut's FindRegions numbers the regions it finds 1..N (it labels its input again); for label images
        # straight from label() that is the label itself, after the label filters it closes the gaps [memory of
        # morphocut/image.py at 03dbc6b, not available offline]
        seq += 1
        if min_intensity is not None and row[F_IMAX] < min_intensity:
            continue
        yield Region(row, shape, padding, labels, image, None if shapes is None else shapes[j], mask_fn,
                     seq if renumber else None, whole, label_fn)


def recalc_metadata(region: Region, meta: Dict, object_id_fmt: Optional[str] = None) -> Dict:
    """loki/pipeline.py:604-619, including its unpacking of ``region.bbox`` as ``(y0, x0, x1, y1)``: skimage's
    bbox is (min_row, min_col, max_row, max_col), so ``object_width`` is max_row - min_col and
    ``object_height`` is max_col - min_row, exactly as the reference computes them."""
    meta = dict(meta)
    (y0, x0, x1, y1) = region.bbox
    meta["object_posx"] = x0
    meta["object_posy"] = y0
    meta["object_sequence"] = region.sequence
    meta["object_width"] = x1 - x0
    meta["object_height"] = y1 - y0
    if object_id_fmt is not None:
        meta["object_id"] = object_id_fmt.format_map(meta)
    # (region.image_intensity[region.image] == 0).mean() -- taken from the exact zero count of the table
    meta["object_frac_invalid"] = float(region[F_FRAC_INVALID])
    return meta


def zooprocess_features(region: Region, meta: Optional[Dict] = None, prefix: str = "object_") -> Dict:
    """``CalculateZooProcessFeatures(region, meta, prefix)`` (loki/pipeline.py:625, 654) from the stage's tables.
    Keys follow morphocut's ZooProcess names.  With a shape row: ``area`` = filled area, ``area_exc`` = pixel
    count, ``%area``, ``perim.``, ``circ.``, ``circex``, ``perimareaexc``, ``perimmajor``, ``euler_number``,
    ``convex_area``, ``solidity`` as well; without one ``area`` falls back to the pixel count and the perimeter-based keys are absent."""
    out = dict(meta) if meta is not None else {}
    row = region._row
    # the reference hands the SAME padded region to recalc_metadata and to CalculateZooProcessFeatures
    # (loki/pipeline.py:604-625): bbox is the padded slice (start clipped, stop not), bbox_area / extent /
    # local_centroid come from region.image, i.e. the slice clipped to the frame
    r0, c0, r1, c1 = region.bbox
    crop_h, crop_w = region.crop_shape
    area = float(row[F_AREA])
    major, minor = float(row[F_AXIS_MAJOR]), float(row[F_AXIS_MINOR])
    mean = float(row[F_IMEAN])
    has_shape = region._shape_row is not None
    filled = region.filled_area if has_shape else area
    bbox_area = float(crop_h * crop_w)
    feats = {
        "label": region.sequence,
        "width": c1 - c0,
        "height": r1 - r0,
        "bx": c0,
        "by": r0,
        "area_exc": area,
        "area": filled,
        "%area": 1.0 - area / filled,
        "major": major,
        "minor": minor,
        "y": float(row[F_CENTROID]),
        "x": float(row[F_CENTROID + 1]),
        "min": float(row[F_IMIN]),
        "max": float(row[F_IMAX]),
        "mean": mean,
        "intden": filled * mean,
        "elongation": (major / minor) if minor > 0 else float("inf"),
        "range": float(row[F_IMAX]) - float(row[F_IMIN]),
        "angle": float(row[F_ORIENT]) / math.pi * 180.0 + 90.0,
        "bounding_box_area": bbox_area,
        "eccentricity": float(row[F_ECC]),
        "equivalent_diameter": math.sqrt(4.0 * area / math.pi),
        "extent": area / bbox_area,
        "local_centroid_row": float(row[F_CENTROID]) - r0,
        "local_centroid_col": float(row[F_CENTROID + 1]) - c0,
    }
    if has_shape:
        perim = region.perimeter
        sq = perim * perim
        feats.update({
            "perim.": perim,
            "circ.": (4.0 * math.pi * filled / sq) if sq > 0 else float("inf"),
            "circex": (4.0 * math.pi * area / sq) if sq > 0 else float("inf"),
            "perimareaexc": perim / area,
            "perimmajor": (perim / major) if major > 0 else float("nan" if perim == 0 else "inf"),
            "euler_number": region.euler_number,
        })
        convex = region.convex_area
        if convex == convex:  # NaN only for objects taller than the hull storage of the kernel
            feats.update({"convex_area": convex, "solidity": area / convex})
    for k, v in feats.items():
        out[prefix + k] = v
    for j in range(7):
        out[f"{prefix}hu{j + 1}"] = float(row[F_HU + j])
    return out


def zooprocess_table(result, padding: int = 75, min_intensity: Optional[float] = None, prefix: str = "object_",
                     whole_frame: bool = False) -> Dict[str, np.ndarray]:
    """The metadata of EVERY object of a batch at once, as columns: what ``objects_of`` (FindRegions ->
    recalc_metadata -> CalculateZooProcessFeatures, loki/pipeline.py:589-625) returns object by object, computed with
    numpy over the whole object table -- one call per batch instead of a Python dict per object (tens of microseconds
    each, which would otherwise be the slowest step behind the GPU).  Column ``image_index`` is the vignette / frame
    of the batch an object belongs to; the rows are in (vignette, label) order; ``object_id`` is left to the caller
    (a format string over its own metadata).  ``whole_frame``: the threshold branch (ImageProperties: the region is
    the whole frame, no padding)."""
    tab = np.asarray(result.table)
    off = np.asarray(result.lab_off).astype(np.int64)
    g = result.geometry
    n_img = len(off) - 1
    img = np.repeat(np.arange(n_img), np.diff(off))
    live = tab[:, F_AREA] > 0
    # FindRegions numbers the regions it finds 1..N per frame (gaps left by the label filters are closed)
    cs = np.cumsum(live)
    before = np.concatenate([[0], cs])[off[:-1]]
    seq = (cs - before[img]).astype(np.int64)
    keep = live.copy()
    if min_intensity is not None and not whole_frame:
        keep &= ~(tab[:, F_IMAX] < min_intensity)
    rows = np.nonzero(keep)[0]
    t = tab[rows]
    iv = img[rows]
    H, W = np.asarray(g.h)[iv].astype(np.int64), np.asarray(g.w)[iv].astype(np.int64)
    b = t[:, F_BBOX:F_BBOX + 4].astype(np.int64)
    if whole_frame:
        r0, c0, r1, c1 = np.zeros_like(H), np.zeros_like(W), H, W
    else:  # start clipped at 0, stop not clipped (morphocut's _enlarge_slice)
        r0, c0 = np.maximum(0, b[:, 0] - padding), np.maximum(0, b[:, 1] - padding)
        r1, c1 = b[:, 2] + padding, b[:, 3] + padding
    crop_h, crop_w = np.minimum(r1, H) - r0, np.minimum(c1, W) - c0
    area = t[:, F_AREA]
    major, minor = t[:, F_AXIS_MAJOR], t[:, F_AXIS_MINOR]
    mean = t[:, F_IMEAN]
    shape = getattr(result, "shape_table", None)
    has_shape = shape is not None
    filled = np.asarray(shape)[rows, S_FILLED_AREA] if has_shape else area
    bbox_area = (crop_h * crop_w).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        cols = {
            "image_index": iv,
            # recalc_metadata (:604-619), including its unpacking of region.bbox as (y0, x0, x1, y1)
            "object_posx": c0, "object_posy": r0, "object_sequence": seq[rows],
            "object_width": r1 - c0, "object_height": c1 - r0,
            "object_frac_invalid": t[:, F_FRAC_INVALID],
        }
        feats = {
            "label": seq[rows], "width": c1 - c0, "height": r1 - r0, "bx": c0, "by": r0,
            "area_exc": area, "area": filled, "%area": 1.0 - area / filled, "major": major, "minor": minor,
            "y": t[:, F_CENTROID], "x": t[:, F_CENTROID + 1], "min": t[:, F_IMIN], "max": t[:, F_IMAX], "mean": mean,
            "intden": filled * mean, "elongation": np.where(minor > 0, major / minor, np.inf),
            "range": t[:, F_IMAX] - t[:, F_IMIN], "angle": t[:, F_ORIENT] / math.pi * 180.0 + 90.0,
            "bounding_box_area": bbox_area, "eccentricity": t[:, F_ECC],
            "equivalent_diameter": np.sqrt(4.0 * area / math.pi), "extent": area / bbox_area,
            "local_centroid_row": t[:, F_CENTROID] - r0, "local_centroid_col": t[:, F_CENTROID + 1] - c0,
        }
        if has_shape:
            s = np.asarray(shape)[rows]
            perim = s[:, S_PERIMETER]
            sq = perim * perim
            convex = s[:, S_CONVEX_AREA]
            feats.update({
                "perim.": perim,
                "circ.": np.where(sq > 0, 4.0 * math.pi * filled / sq, np.inf),
                "circex": np.where(sq > 0, 4.0 * math.pi * area / sq, np.inf),
                "perimareaexc": perim / area,
                "perimmajor": np.where(major > 0, perim / major, np.where(perim == 0, np.nan, np.inf)),
                "euler_number": s[:, S_EULER].astype(np.int64),
                "convex_area": convex,          # NaN for objects taller than the hull storage of the kernel
                "solidity": area / convex,
            })
    for k, v in feats.items():
        cols[prefix + k] = v
    for j in range(7):
        cols[f"{prefix}hu{j + 1}"] = t[:, F_HU + j]
    return cols


def extract_roi(image: np.ndarray, region: Region, alpha: float = 0, bg_color=0, keep_background: bool = False) -> np.ndarray:
    """``ExtractROI(image, region, alpha=1 if config.apply_mask else 0, bg_color=config.background_color,
    keep_background=config.keep_background)`` (loki/pipeline.py:596-602): the padded crop ``image[region.slice]``.
    ``apply_mask`` defaults to False (config_schema.py:97-100), i.e. the plain crop.  With ``alpha`` = 1 the pixels
    that are "not part of the current object" (config_schema.py:97-100) are painted with the scalar ``bg_color``:
    everything outside the object -- or, with ``keep_background`` ("when hiding non-object image regions, keep
    background", config_schema.py:105-107), only the pixels of OTHER objects, the unlabelled background stays.
    Restated from the schema's descriptions: morphocut's ExtractROI is not available offline (parity unpinned); colour
    names / quantiles of ``background_color`` and fractional ``alpha`` raise NotImplementedError."""
    crop = np.asarray(image)[region.slice]
    if alpha == 0:
        return crop
    if alpha != 1 or not np.isscalar(bg_color):
        raise NotImplementedError("only alpha in {0, 1} with a scalar background colour")
    if keep_background:
        lab = region.label_image
        keep = (lab == 0) | (lab == region.label)
    else:
        keep = region.image
    return np.where(keep, crop, np.asarray(bg_color, dtype=crop.dtype))


def objects_of(result, i: int, meta: Optional[Dict] = None, padding: int = 75, min_intensity: Optional[float] = None,
               image: Optional[np.ndarray] = None, object_id_fmt: Optional[str] = None) -> List[Dict]:
    """All objects of vignette / frame ``i`` as metadata dicts: FindRegions -> recalc_metadata ->
    CalculateZooProcessFeatures, in the order of loki/pipeline.py:589-625 (default padding 75,
    loki/config_schema.py:90-96)."""
    base = {} if meta is None else meta
    return [zooprocess_features(r, recalc_metadata(r, base, object_id_fmt))
            for r in find_regions(result, i, padding, min_intensity, image)]
