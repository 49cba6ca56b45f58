"""The batched LOKI re-segmentation stage.

Takes the two EXISTING config objects of the reference as its parameters
(``ThresholdSegmentationConfig``, ``SegmentationPostprocessingConfig``,
maze_ipp/loki/config_schema.py:8-37 -- any object with the same attributes works) and runs, for a
whole batch of vignettes at once and entirely on the GPU, the chain the reference runs per vignette:

* ``threshold`` only (loki/pipeline.py:648-656): ``mask = image > threshold_brighter``; vignettes
  with an empty mask are dropped; the whole mask is ONE region (``ImageProperties``).
* ``postprocess`` only (loki/pipeline.py:396-459, 589-625): bool cast of the model's foreground
  prediction, opening, closing, label, clear_border, remove_small_objects, merge_labels, then one
  region per surviving label (``FindRegions``).
* both (benchmark configuration): threshold feeding the post-processing chain.

Opening runs BEFORE closing, as in the reference.  Morphology is the EDT-based
``maze_ipp/isotropic.py`` (the north-star contract), not the live ``disk(r, "crosses")`` footprints.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

import ctypes
import os
import threading

from ._lib import (BAND_SINGLE_REGION, MAZE_ERR_TYPEERROR, NACC, NEXT, NFEAT, NSHAPE, RP_HIGH_ORDER, STEP_COMPACT, StepArgs,
                   check, lib)
from ._lib import MAX_DISK_RADIUS
from .device import (Arena, BatchGeometry, DeviceBatch, fold_dilation_radius, fold_erosion_radius, fold_threshold)


@dataclass
class ThresholdSegmentationConfig:
    """Mirror of maze_ipp/loki/config_schema.py:32-37."""
    threshold_brighter: float


@dataclass
class SegmentationPostprocessingConfig:
    """Mirror of maze_ipp/loki/config_schema.py:8-29 (same names, same defaults)."""
    closing_radius: int = 0
    opening_radius: int = 0
    merge_segments_distance: int = 0
    min_area: int = 0
    n_threads: int = 0
    clear_border: bool = False


class DeviceResult:
    """Outputs of one batch, still on the device.  On the asynchronous fused path the object counts are
    read back lazily: ``finalize()`` (called by every accessor that needs them) waits for the readback
    and, in the rare case that a vignette overflowed the fused kernel's tables, redoes the batch with the
    per-operator kernels."""

    def __init__(self, batch, bits, labels, lab_off, table, n_obj, keep=None, merge_status=None, mask=None,
                 pending=None):
        self.mask = mask
        self.batch = batch
        self.bits = bits
        self.labels = labels
        self.lab_off = lab_off
        self._table = table
        self._n_obj = n_obj
        self.keep = keep
        self.merge_status = merge_status
        self._pending = pending
        # True when the table was produced by kernels on the main stream that no event has covered yet
        # (every path except the asynchronous fused one, and the redo inside finalize)
        self._sync_main = pending is None
        self.redone = False
        self.ready = None  # event after which mask / labels / bits are complete (None: stream order of the caller)
        # band pipeline: run list (maze_run_t as int64) + per-band {base, n_runs, ..}, host copies of the band plan,
        # runs used (after finalize), vignettes whose outputs exist only as dense arrays
        self.runs = self.band_out = self.bands_host = self.band_off_host = None
        self.n_runs = 0
        self.compact = False
        self.dense_only = []
        self.single_region = False  # threshold branch on the band pipeline: one region per vignette, no label image

    def finalize(self):
        if self._pending is not None:
            pending, self._pending = self._pending, None
            pending(self)
        return self

    @property
    def n_obj(self):
        return self.finalize()._n_obj

    @property
    def table(self):
        return self.finalize()._table


class _HostArrayPool:
    """Flat host arrays for StageResult.materialize().  A fresh ``np.empty`` of a gigabyte is mapped in page by page
    (and zeroed by the kernel) while the expansion writes it, every batch; an array that nothing refers to any more --
    no StageResult, no per-vignette view: checked by its reference count -- is handed out again instead."""

    def __init__(self, keep=6):
        self._bufs = []
        self._keep = keep
        self._lock = threading.Lock()

    def take(self, n, dtype):
        import sys
        dtype = np.dtype(dtype)
        with self._lock:
            for b in self._bufs:
                # list entry + loop variable + getrefcount's argument: nobody else holds it (views keep their base alive)
                if b.dtype == dtype and n <= b.size <= 2 * n + 4096 and sys.getrefcount(b) == 3:
                    return b[:n]
            b = np.empty(max(int(n), 1), dtype)
            idle = [k for k, o in enumerate(self._bufs) if sys.getrefcount(o) == 3]
            if len(self._bufs) >= self._keep and idle:
                del self._bufs[idle[0]]
            if len(self._bufs) < self._keep:
                self._bufs.append(b)
            return b[:n]


_HOST_POOL = _HostArrayPool()


class StageResult:
    """Host-side outputs of one batch: per-vignette mask / label image and the object table.

    Dense form: flat mask / label buffers (numpy views per vignette).  Compact form (LokiSegmentationStage(compact=
    True)): the label image of every vignette is held as its RUN LIST {y, x0, x1, label} -- the form in which it
    crossed PCIe, 8 bytes per run instead of 5 bytes per pixel -- and ``mask(i)`` / ``labels(i)`` / ``object_mask``
    expand it on demand (native, include/maze_b200.h: maze_host_expand / maze_host_expand_crop); ``materialize()``
    expands the whole batch at once.  The arrays are what the reference's stage returns (bool mask, int32 labels,
    loki/pipeline.py:459) either way."""

    def __init__(self, geometry: BatchGeometry, mask_flat, labels_flat, lab_off, table, keep=None):
        self.geometry = geometry
        self._mask = mask_flat
        self._labels = labels_flat
        self.lab_off = lab_off
        self.table = table
        self.keep = keep  # threshold branch: vignettes that survive the empty-mask filter
        self.shape_table = None  # shape_features=True: (n_obj, NSHAPE) perimeter / filled_area / euler_number rows
        self.merge_failed = None  # merge_errors="ignore": vignettes where merge_labels hit the reference's TypeError
        self._no_labels = False  # threshold branch: there is no label image (labels(i) is None)
        # compact form
        self._runs = self._band_out = self._band_off = self._rpb = None
        self._dense = {}    # vignettes without a run list (per-operator kernels): i -> (mask, labels)
        self._cache = {}    # vignettes expanded on demand

    @classmethod
    def from_runs(cls, geometry, runs, band_out, band_off, rpb, dense, lab_off, table):
        out = cls(geometry, None, None, lab_off, table)
        out._runs, out._band_out, out._band_off, out._rpb, out._dense = runs, band_out, band_off, rpb, dense
        return out

    @property
    def compact(self) -> bool:
        return self._runs is not None

    def __len__(self):
        return self.geometry.n_img

    def _expand(self, i):
        if i in self._dense:
            m, l = self._dense[i]
            if m is None:  # merge_labels rewrote the label image (it came down dense); the mask is still in the runs
                m = self._expand_runs(i)[0]
                self._dense[i] = (m, l)
            return self._dense[i]
        return self._expand_runs(i)

    def _expand_runs(self, i):
        if i not in self._cache:
            g = self.geometry
            h, w = int(g.h[i]), int(g.w[i])
            mask = np.empty((h, w), np.uint8)
            labels = np.empty((h, w), np.int32)
            lo, hi = np.asarray([self._band_off[i]], np.int32), np.asarray([self._band_off[i + 1]], np.int32)
            hh, ww = np.asarray([h], np.int32), np.asarray([w], np.int32)
            pm = np.asarray([mask.ctypes.data], np.uint64)
            pl = np.asarray([labels.ctypes.data], np.uint64)
            check(lib().maze_host_expand(self._runs.ctypes.data, self._band_out.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                         hh.ctypes.data, ww.ctypes.data, 1, pm.ctypes.data, pl.ctypes.data, 1),
                  "maze_host_expand")
            if len(self._cache) >= 64:
                self._cache.clear()
            self._cache[i] = (mask, labels)
        return self._cache[i]

    def mask(self, i) -> np.ndarray:
        if self._runs is not None:
            return self._expand(i)[0].view(bool)
        return self.geometry.view(self._mask, i).view(bool)

    def labels(self, i) -> Optional[np.ndarray]:
        if self._no_labels:
            return None
        if self._runs is not None:
            return self._expand(i)[1]
        return None if self._labels is None else self.geometry.view(self._labels, i)

    def runs(self, i) -> np.ndarray:
        """Run list of vignette i (compact form): structured array with fields y, x0, x1 (inclusive), label."""
        if self._runs is None:
            raise ValueError("dense result: no run list")
        if i in self._dense:
            raise ValueError(f"vignette {i} has no run list of its label image (dense arrays: labels(i))")
        bo = self._band_out[int(self._band_off[i]):int(self._band_off[i + 1])]
        parts = [self._runs[int(b["base"]):int(b["base"]) + int(b["n_runs"])] for b in bo if b["base"] >= 0]
        return np.concatenate(parts) if parts else self._runs[:0]

    def object_mask(self, i, sl, label=None) -> np.ndarray:
        """Boolean crop ``labels(i)[sl] == label`` (``label=None``: ``mask(i)[sl]``) without expanding the vignette:
        what RegionProperties.image / ExtractROI read (loki/pipeline.py:589-602)."""
        g = self.geometry
        h, w = int(g.h[i]), int(g.w[i])
        r0, r1, _ = sl[0].indices(h)
        c0, c1, _ = sl[1].indices(w)
        if self._runs is None or i in self._dense:
            lab = self.labels(i)
            return (lab[r0:r1, c0:c1] == label) if label is not None else self.mask(i)[r0:r1, c0:c1]
        out = np.empty((max(r1 - r0, 0), max(c1 - c0, 0)), np.uint8)
        check(lib().maze_host_expand_crop(self._runs.ctypes.data, self._band_out.ctypes.data, int(self._band_off[i]),
                                          int(self._band_off[i + 1]), int(self._rpb[i]), r0, max(r1, r0), c0, max(c1, c0),
                                          0 if label is None else int(label), out.ctypes.data, None),
              "maze_host_expand_crop")
        return out.view(bool)

    def label_crop(self, i, sl) -> np.ndarray:
        """``labels(i)[sl]`` (int32) without expanding the vignette: the crop in which ExtractROI tells the object from
        the others (loki/pipeline.py:596-602)."""
        g = self.geometry
        h, w = int(g.h[i]), int(g.w[i])
        r0, r1, _ = sl[0].indices(h)
        c0, c1, _ = sl[1].indices(w)
        if self._runs is None or i in self._dense:
            lab = self.labels(i)
            if lab is None:
                raise ValueError("this result has no label image (threshold branch)")
            return lab[r0:r1, c0:c1]
        out = np.empty((max(r1 - r0, 0), max(c1 - c0, 0)), np.int32)
        check(lib().maze_host_expand_crop(self._runs.ctypes.data, self._band_out.ctypes.data, int(self._band_off[i]),
                                          int(self._band_off[i + 1]), int(self._rpb[i]), r0, max(r1, r0), c0, max(c1, c0),
                                          0, None, out.ctypes.data), "maze_host_expand_crop")
        return out

    def materialize(self, threads=None) -> "StageResult":
        """Compact -> dense: every mask and label image of the batch as flat host arrays (multi-threaded)."""
        if self._runs is None:
            return self
        g = self.geometry
        n = g.n_img
        mask = _HOST_POOL.take(g.total_px, np.uint8)
        labels = _HOST_POOL.take(g.total_px, np.int32)
        if threads is None and os.environ.get("MAZE_EXPAND_THREADS"):
            threads = int(os.environ["MAZE_EXPAND_THREADS"])
        if threads is None:
            try:
                avail = len(os.sched_getaffinity(0))
                lws = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))  # ranks of one box share its cores
                threads = min(16, max(1, avail // 2)) if lws == 1 else min(8, max(2, avail // lws))
            except Exception:
                threads = 4
        todo = np.asarray([i for i in range(n) if i not in self._dense or self._dense[i][0] is None], np.int64)
        if len(todo):
            pm = (mask.ctypes.data + g.pix_off[todo]).astype(np.uint64)
            pl = (labels.ctypes.data + 4 * g.pix_off[todo]).astype(np.uint64)
            lo = np.ascontiguousarray(self._band_off[todo], np.int32)
            hi = np.ascontiguousarray(self._band_off[todo + 1], np.int32)
            hh = np.ascontiguousarray(g.h[todo], np.int32)
            ww = np.ascontiguousarray(g.w[todo], np.int32)
            check(lib().maze_host_expand(self._runs.ctypes.data, self._band_out.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                         hh.ctypes.data, ww.ctypes.data, len(todo), pm.ctypes.data, pl.ctypes.data,
                                         int(threads)), "maze_host_expand")
        for i, (m, l) in self._dense.items():
            if m is not None:
                g.view(mask, i)[...] = m
            if l is not None:
                g.view(labels, i)[...] = l
        self._mask, self._labels = mask, labels
        self._runs = self._band_out = None
        self._dense, self._cache = {}, {}
        return self

    def features(self, i) -> np.ndarray:
        """Rows of the object table that belong to vignette i (row k = label k + 1)."""
        return self.table[int(self.lab_off[i]):int(self.lab_off[i + 1])]

    def shape_features(self, i) -> Optional[np.ndarray]:
        """Rows of the shape table (perimeter, filled_area, euler_number, ...) that belong to vignette i."""
        if self.shape_table is None:
            return None
        return self.shape_table[int(self.lab_off[i]):int(self.lab_off[i + 1])]


class Workspace:
    """Device buffers of the fused path, allocated once and grown on demand so that the steady state makes
    no allocator calls.  The stage rotates n_lanes (default six) of them (each with its own lane / side stream and scratch
    arena); results of run_device alias one and stay valid for the next n_lanes - 1 run_device calls."""

    def __init__(self):
        self._t = {}

    def get(self, key, n, dtype, device):
        t = self._t.get(key)
        if t is None or t.numel() < n or t.device != device:
            t = torch.empty(int(n * 1.25) + 16, dtype=dtype, device=device)
            self._t[key] = t
        return t[:n]


class _PinnedPool:
    """Reusable pinned host staging buffers and device input buffers (one per purpose), grown on demand."""

    def __init__(self):
        self._bufs = {}
        self._dev = {}

    def dev(self, key, n, dtype, device):
        """Device buffer for an upload: a fresh torch allocation per batch costs milliseconds (cudaMalloc)."""
        t = self._dev.get(key)
        if t is None or t.numel() < n or t.dtype != dtype or t.device != device:
            t = torch.empty(int(max(n, 1) * 1.3) + 64, dtype=dtype, device=device)
            self._dev[key] = t
        return t[:n]

    def get(self, key, n, dtype):
        t = self._bufs.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(int(max(n, 1) * 1.3) + 64, dtype=dtype, pin_memory=True)  # pinning is slow: grow rarely
            self._bufs[key] = t
        return t[:n]


class LokiSegmentationStage:
    def __init__(self, threshold=None, postprocess=None, device=None, high_order=True, fused=True,
                 merge_errors="raise", shape_features=False, morphology="isotropic", pipeline=None, compact=False,
                 n_lanes=None):
        """pipeline: "bands" (default; maze_band_stage: band front, run-list labelling, dense writer) or "fused" (the
        vignette-resident kernel maze_vignette_stage); MAZE_PIPELINE overrides the default.
        compact: the per-pixel outputs stay on the device as the RUN LIST {y, x0, x1, label} (8 bytes per run, a
        few thousand runs per megapixel) and cross PCIe in that form; StageResult.mask(i) / labels(i) / crops expand
        them on the host on demand.  Needs the band pipeline without label filters / merge (otherwise the dense
        arrays are downloaded as before).
        morphology: "isotropic" (maze_ipp/isotropic.py, the EDT-based operators of the north-star contract) or
        "crosses" (what the live pipeline calls: skimage binary_opening / binary_closing with
        disk(radius, decomposition="crosses"), loki/pipeline.py:408-427).
        shape_features: also produce, per object, the RegionProperties values CalculateZooProcessFeatures reads
        besides the moments -- perimeter, filled_area, euler_number (StageResult.shape_table; one more kernel over
        the label image per batch).
        merge_errors: "raise" (the reference's behaviour: merge_labels raises TypeError when a bridge
        swallows a label, merge_labels.py:19-20, and the run aborts) or "ignore" (keep the labels as the loop
        left them for those vignettes and list them in StageResult.merge_failed)."""
        self.fused = fused
        self.pipeline = pipeline or os.environ.get("MAZE_PIPELINE", "bands")
        if self.pipeline not in ("bands", "fused"):
            raise ValueError("pipeline must be 'bands' or 'fused'")
        self.compact = bool(compact)
        self.merge_errors = merge_errors
        self.shape_features = shape_features
        if morphology not in ("isotropic", "crosses"):
            raise ValueError("morphology must be 'isotropic' or 'crosses'")
        self.morphology = morphology
        if threshold is None and postprocess is None:
            raise ValueError("exactly one of threshold / postprocess (or both, for the composite stage) is required")
        self.threshold = threshold
        self.postprocess = postprocess
        self.device = device
        self.high_order = high_order
        self._pool = _PinnedPool()
        self._ws = Workspace()
        # workspaces (each with its own lane stream) in rotation; map() keeps batch i+1 in flight while batch i is
        # downloaded, so fewer than two would let a batch overwrite the buffers its predecessor is still read from
        self.n_lanes = int(n_lanes if n_lanes is not None else os.environ.get("MAZE_LANES", "6"))
        if self.n_lanes < 1:
            raise ValueError("MAZE_LANES must be >= 1")
        if shape_features:
            self.compact = False  # maze_label_shape reads the dense label image on the device
        self._ws_ring, self._ws_i = [Workspace() for _ in range(self.n_lanes)], 0
        self._copy_stream = None
        self._small_copy_stream = None
        self._map_pools = [_PinnedPool(), _PinnedPool(), _PinnedPool()]
        self._img_ring = None  # pinned staging buffers of map() (four sets, allocated on first use)
        self._shared_busy = None  # download event of the last batch that used the shared (non-rotating) workspace
        self._merge_busy = None   # last merge_labels launch that uses the shared distance maps (deferred merge tails)
        self._readback, self._readback_i = [], 0
        self._readback_graph = []
        self.graphs = os.environ.get("MAZE_GRAPHS", "1") != "0"  # CUDA graphs for steps of a few vignettes (frames)

    # ---- device-resident core ----------------------------------------------------------------------
    def _passes(self):
        """[(d2 threshold, invert)] of the morphology passes, or None when a radius needs the exact-EDT path
        (cached per configuration: the footprint tables of the crosses mode take milliseconds to build)."""
        pp = self.postprocess
        key = (self.morphology, pp.opening_radius, pp.closing_radius)
        cache = self.__dict__.setdefault("_passes_cache", {})
        if key not in cache:
            cache[key] = self._passes_uncached()
        return cache[key]

    def _passes_uncached(self):
        pp = self.postprocess
        out = []
        if self.morphology == "crosses":  # loki/pipeline.py:408-427: erosion + dilation / dilation + erosion
            from .morphology import disk, footprint_pass_code
            if max(pp.opening_radius, pp.closing_radius) > MAX_DISK_RADIUS:
                raise NotImplementedError(f"footprint radii up to {MAX_DISK_RADIUS}")
            if pp.opening_radius > 0:
                code = footprint_pass_code(disk(int(pp.opening_radius), decomposition="crosses"))
                out += [(code, 0), (code, 1)]
            if pp.closing_radius > 0:
                code = footprint_pass_code(disk(int(pp.closing_radius), decomposition="crosses"))
                out += [(code, 1), (code, 0)]
            return out
        if pp.opening_radius > 0:   # isotropic.py:97-98
            out += [(fold_erosion_radius(pp.opening_radius), 0), (fold_dilation_radius(pp.opening_radius), 1)]
        if pp.closing_radius > 0:   # isotropic.py:128-129
            out += [(fold_dilation_radius(pp.closing_radius), 1), (fold_erosion_radius(pp.closing_radius), 0)]
        if any(t >= (MAX_DISK_RADIUS + 1) ** 2 for t, _ in out):
            return None
        # a pass with d2 threshold 0 is the identity for both compares (d2 > 0 and d2 <= 0 both mean
        # "is foreground"; the phantom pixel lies outside the image): e.g. the dilation of opening r=1
        return [(t, inv) for t, inv in out if t != 0]

    def _front_generic(self, batch, d_src, t_int, labels=None, mask=None):
        """Per-operator kernels: threshold -> opening -> closing -> label (any size, any radius)."""
        pp = self.postprocess
        bits, flags = batch.threshold_pack(d_src, t_int)
        if self.morphology == "crosses":
            for code, inv in self._passes():
                bits, flags = batch.morph_pass(bits, flags, code, inv)
        else:
            if pp.opening_radius > 0:
                bits, flags = batch.opening(bits, flags, pp.opening_radius)
            if pp.closing_radius > 0:
                bits, flags = batch.closing(bits, flags, pp.closing_radius)
        labels, lab_off = batch.label(bits, labels=labels)
        mask = batch.unpack_mask(bits, out=mask)
        return bits, labels, lab_off, mask

    def run_device(self, batch: DeviceBatch, d_image, d_pred=None) -> DeviceResult:
        """All kernels for one resident batch.  d_image: flat uint8 intensities; d_pred: flat uint8
        foreground prediction (postprocess-only mode)."""
        pp = self.postprocess
        g = batch.g
        if self.threshold is not None:
            d_src, t_int = d_image, fold_threshold(self.threshold.threshold_brighter)
        else:
            d_src, t_int = d_pred, 0  # np.asarray(pred, dtype=bool), loki/pipeline.py:405
        if pp is None and self._threshold_bands(batch):
            # ImageProperties(mask, image) through the band pipeline: threshold, run list, ONE region per vignette
            return self._run_fused_async(batch, d_src, d_image, t_int, [])
        if pp is None:
            return self.run_device_threshold_generic(batch, d_src, d_image, t_int)
        passes = self._passes() if self.fused else None
        filters = pp.clear_border or pp.min_area > 0 or pp.merge_segments_distance > 0
        if passes is not None and (not filters or self._band_filters(batch)):
            return self._run_fused_async(batch, d_src, d_image, t_int, passes)
        if passes is not None and pp.merge_segments_distance > 0 and self._band_filters(batch, merge_ok=True):
            return self._run_merge_on_bands(batch, d_src, d_image, t_int, passes)
        return self._run_filter_path(batch, d_src, d_image, t_int, passes)

    def _run_merge_on_bands(self, batch, d_src, d_image, t_int, passes):
        """merge_labels behind the band pipeline: the band step (labels, label filters, table) runs first; the object
        table then tells, EXACTLY, which vignettes merge_labels can change at all, and only those go through the
        merge kernel; the rows of the vignettes in which labels did merge are recomputed.  The tail (candidates, merge,
        rows) runs in DeviceResult.finalize().

        The reference's loop (merge_labels.py:81-96) grows ONE cluster from the smallest label l0 and stops at the first
        candidate whose merge distance exceeds max_distance.  If every other label l lies farther than max_distance from
        l0 -- certain when the bounding boxes are more than max_distance apart in rows or in columns -- then whichever
        candidate the first iteration picks, min(distmap + cur_distmap) > max_distance: inside both windows the sum is
        at least dist(l0, l) (triangle inequality, the windowed EDTs are exact there), and outside a window the map
        holds its window maximum, which is at least pad = ceil(max_distance) + 1 as soon as the window reaches one
        unclipped margin.  The loop then breaks at once and the aliased call returns the labels unchanged."""
        pp = self.postprocess
        g = batch.g
        n = g.n_img
        saved_compact, self.compact = self.compact, False  # merge_labels needs the dense label image on the device
        try:
            res = self._run_fused_async(batch, d_src, d_image, t_int, passes)
        finally:
            self.compact = saved_compact
        lane = self._ws_ring[(self._ws_i - 1) % self.n_lanes].lane
        band_pending, res._pending = res._pending, None

        def pending(r):
            # the merge tail runs when the result is finalised, on the batch's own lane stream: the band steps of the
            # batches behind it (other lanes) fill the SMs that the long vignettes of the merge kernel leave idle
            band_pending(r)
            if r.redone or len(r.dense_only):
                # flagged vignettes: the plain filter path takes the whole batch
                r2 = self._run_filter_path(batch, d_src, d_image, t_int, passes)
                r.redone = True
                r.dense_only = list(range(n))
                r.bits, r.labels, r.mask, r.lab_off = r2.bits, r2.labels, r2.mask, r2.lab_off
                r._table, r._n_obj, r.merge_status = r2._table, r2._n_obj, r2.merge_status
                r.runs = None
                r._sync_main = True
                r.ready = torch.cuda.Event()
                r.ready.record(torch.cuda.current_stream())
                return
            with torch.cuda.stream(lane):
                tab = r._table[:, :6].cpu().numpy()          # label, area, bbox
                off = r.lab_off.cpu().numpy().astype(np.int64)
                need = _merge_candidates(tab, off, g.h, g.w, float(pp.merge_segments_distance))
                merge_status = torch.zeros(n, dtype=torch.int32, device=batch.device)
                if len(need):
                    n_obj = int(off[-1])
                    if self._merge_busy is not None:
                        lane.wait_event(self._merge_busy)  # ONE set of distance maps for all lanes
                    scratch = tuple(self._ws.get(k, g.total_px, torch.int32, batch.device) for k in ("d2a", "d2b", "d2c"))
                    saved_arena, batch.arena = batch.arena, None  # (scratch of this tail does not belong to a lane)
                    _, n_merge, _, merge_status, _ = batch.merge_labels(r.labels, r.labels, r.lab_off, n_obj,
                                                                        pp.merge_segments_distance, only=need, scratch=scratch)
                    self._merge_busy = torch.cuda.Event()
                    self._merge_busy.record(lane)
                    # rows of the vignettes in which something merged, from the dense label image (bridges leave the
                    # runs of the bit plane).  Which ones is decided on the device -- no round trip: the tiles of every
                    # candidate are launched, those of a vignette without a merge return at once
                    acc_base = torch.where(n_merge > 0, -1, 0).to(torch.int32)
                    cap = r._table.shape[0]
                    batch.regionprops(r.lab_off, cap, labels=r.labels, bits=None, image=d_image,
                                      high_order=self.high_order, runs=False, table=r._table, acc_base=acc_base,
                                      tiles=batch.tiles_of(need))
                    batch.arena = saved_arena
                r.merge_status = merge_status
                r.n_merge = n_merge if len(need) else torch.zeros(n, dtype=torch.int32, device=batch.device)
                r.compact = saved_compact  # (the step ran dense on the device; the TRANSPORT can still be compact)
                r._sync_main = True
                r.ready = torch.cuda.Event()
                r.ready.record(lane)

        res._pending = pending
        res.deferred_pixels = True  # the label image is final only after finalize()
        return res

    def run_device_threshold_generic(self, batch, d_src, d_image, t_int) -> DeviceResult:
        """Threshold branch with the per-operator kernels: ImageProperties(mask, image), one region per vignette; empty
        masks are dropped (loki/pipeline.py:651-653)."""
        g = batch.g
        if self._shared_busy is not None:
            self._shared_busy.synchronize()
            self._shared_busy = None
        bits, flags = batch.threshold_pack(d_src, t_int)
        lab_off, n_obj = batch.lab_off_from_bounds(np.ones(g.n_img, np.int64))
        table = batch.regionprops(lab_off, n_obj, bits=bits, image=d_image, high_order=self.high_order)
        keep = (flags & 1).bool()
        return DeviceResult(batch, bits, None, lab_off, table, n_obj, keep=keep, mask=batch.unpack_mask(bits))

    def _threshold_bands(self, batch) -> bool:
        """The threshold branch (loki/pipeline.py:648-656) runs on the band pipeline unless the batch holds frames or
        vignettes a band cannot take."""
        if self.pipeline != "bands" or not self.fused:
            return False
        from ._lib import HUGE_PX
        if (batch.g.npx >= HUGE_PX).any():
            return False
        return len(batch.band_lists(0)[3]) == 0

    def _band_filters(self, batch, merge_ok=False) -> bool:
        """clear_border / remove_small_objects run on the run list inside the band pipeline (no merge_labels, no frames,
        no vignette that needs the per-operator kernels)."""
        pp = self.postprocess
        if self.pipeline != "bands" or (pp.merge_segments_distance > 0 and not merge_ok):
            return False
        from ._lib import HUGE_PX
        from .morphology import pass_radius
        if (batch.g.npx >= HUGE_PX).any():
            return False
        return len(batch.band_lists(sum(pass_radius(t) for t, _ in self._passes()))[3]) == 0

    def _run_filter_path(self, batch, d_src, d_image, t_int, passes) -> DeviceResult:
        """Label filters / merge_labels with the per-operator kernels on the dense label image (synchronous, ONE shared
        workspace): the vignette-resident kernel for the labels, then clear_border / remove_small_objects /
        merge_labels, then regionprops."""
        if self._merge_busy is not None:
            self._merge_busy.synchronize()
            self._merge_busy = None
        pp = self.postprocess
        g = batch.g
        filters = pp.clear_border or pp.min_area > 0 or pp.merge_segments_distance > 0
        if self._shared_busy is not None:  # a download of the previous batch may still read the shared workspace
            self._shared_busy.synchronize()
            self._shared_busy = None
        live_counts = None
        if passes is None:
            bits, labels, lab_off, mask = self._front_generic(batch, d_src, t_int)
            n_obj = int(lab_off[-1].item())  # one 4-byte readback sizes the object table
        else:
            # fused kernel for the labels only (the filters change them before the features are taken)
            n = g.n_img
            ws, dev = self._ws, batch.device
            bits = ws.get("bits", max(g.total_words, 1), torch.int32, dev)
            mask = ws.get("mask", g.total_px, torch.uint8, dev)
            labels = ws.get("labels", g.total_px, torch.int32, dev)
            counts = ws.get("counts", 3 * n, torch.int32, dev)
            counts[:2 * n].zero_()
            counts[2 * n:].fill_(-1)
            n_labels = counts[:n]
            staging = tuple(ws.get(k, 8, dt, dev) for k, dt in (("acc", torch.int64), ("hi", torch.float64),
                                                                 ("ext", torch.int32), ("counter", torch.int32)))
            left = batch.vignette_stage(d_src, d_image, t_int, passes, bits, mask, labels, counts, staging, 0,
                                        high_order=self.high_order, props=False)
            if len(left):
                self._redo_generic(batch, left, d_src, t_int, bits, mask, labels, n_labels)
            h_counts = counts.cpu().numpy()
            redo = np.nonzero(h_counts[n:2 * n])[0]
            if len(redo):
                self._redo_generic(batch, redo, d_src, t_int, bits, mask, labels, n_labels)
                h_counts = counts.cpu().numpy()
            lab_off, n_obj = batch.lab_off_from_bounds(h_counts[:n])
            live_counts = h_counts[:n]
        merge_status = None
        if n_obj > 0 and filters:
            if pp.clear_border:
                batch.clear_border(labels, lab_off, n_obj)
            if pp.min_area > 0:
                batch.remove_small_objects(labels, lab_off, n_obj, pp.min_area)
            if pp.merge_segments_distance > 0:
                merge_status = batch.merge_labels(labels, labels, lab_off, n_obj, pp.merge_segments_distance,
                                                  live=live_counts)[3]
        # merge_labels paints bridges over background, so only then do labels leave the runs of `bits`
        runs = merge_status is None
        table = batch.regionprops(lab_off, n_obj, labels=labels, bits=bits if runs else None, image=d_image,
                                  high_order=self.high_order, runs=runs)
        return DeviceResult(batch, bits, labels, lab_off, table, n_obj, merge_status=merge_status, mask=mask)

    def _run_fused_async(self, batch, d_src, d_image, t_int, passes) -> DeviceResult:
        """n_lanes workspaces (default six), each with its own LANE stream, rotate: batch i+1 starts on the next lane while
        the tail of batch i (label offsets, feature rows, stragglers of the big size class) still runs, so the
        GPU never drains between batches.  The caller's stream is joined at the start (inputs) only; the
        result carries a `ready` event and DeviceResult.finalize() waits for it."""
        dev = batch.device
        ws = self._ws_ring[self._ws_i % self.n_lanes]   # results stay valid for the next n_lanes - 1 calls
        ws.index = self._ws_i % self.n_lanes
        self._ws_i += 1
        if getattr(ws, "lane", None) is None or ws.lane.device != dev:
            ws.lane = torch.cuda.Stream(device=dev)
            ws.side = torch.cuda.Stream(device=dev)
        caller = torch.cuda.current_stream()
        ws.lane.wait_stream(caller)
        with torch.cuda.stream(ws.lane):
            res = self._run_fused_lane(batch, d_src, d_image, t_int, passes, ws)
        return res

    def join(self):
        """Make the current stream wait for everything the stage has in flight on its lanes."""
        cur = torch.cuda.current_stream()
        for ws in self._ws_ring:
            if getattr(ws, "lane", None) is not None:
                cur.wait_stream(ws.lane)
                cur.wait_stream(ws.side)

    def _run_fused_lane(self, batch, d_src, d_image, t_int, passes, ws) -> DeviceResult:
        """threshold -> morphology -> label -> regionprops with no host synchronisation: the
        vignette-resident kernel on the current stream, the few vignettes it cannot hold through the
        per-operator kernels on a forked stream, label offsets scanned on the device, counts read back
        asynchronously (DeviceResult.finalize)."""
        g, dev = batch.g, batch.device
        n = g.n_img
        if getattr(ws, "arena", None) is None or ws.arena.device != dev:
            ws.arena = Arena(dev)
        ws.arena.reset()
        batch.arena = ws.arena
        main = torch.cuda.current_stream()  # == the lane stream of this workspace
        side = ws.side
        if getattr(ws, "side_done", None) is not None:
            main.wait_event(ws.side_done)  # trailing side-stream work of the batch that used this workspace
            ws.side_done = None
        # a step of a few vignettes is bound by the HOST (buffers, descriptor block, band plan of this function: ~0.1 ms):
        # what the function works out is kept per lane and reused while batch, inputs and parameters stay the same
        ppk = None if self.postprocess is None else (bool(self.postprocess.clear_border), int(self.postprocess.min_area))
        ck = (d_src.data_ptr(), d_image.data_ptr(), int(t_int), tuple(passes), self.compact, self.high_order, self.pipeline,
              self.graphs, ppk)
        plan = getattr(ws, "plan", None)
        if plan is not None and plan[1] is batch and plan[0] == ck:
            (bits, mask, labels, counts, lab_off, n_labels, acc_base, cap, staging, table, use_bands, bands_h, band_off_h,
             runs, band_out, left, host, a, keep, huge_pairs, entry, single, use_graph) = plan[2]
        else:
            bits = ws.get("bits", max(g.total_words, 1), torch.int32, dev)
            mask = ws.get("mask", g.total_px, torch.uint8, dev)
            labels = ws.get("labels", g.total_px, torch.int32, dev)
            counts = ws.get("counts", 3 * n, torch.int32, dev)
            lab_off = ws.get("lab_off", n + 1, torch.int32, dev)
            n_labels, acc_base = counts[:n], counts[2 * n:]
            cap = _stage_cap(g)
            staging = (ws.get("acc", cap * NACC, torch.int64, dev), ws.get("hi", cap * 8, torch.float64, dev),
                       ws.get("ext", cap * NEXT, torch.int32, dev), ws.get("counter", 1, torch.int32, dev))
            table = ws.get("table", cap * NFEAT, torch.float64, dev).view(cap, NFEAT)
            use_bands = self.pipeline == "bands"
            bands_h = band_off_h = None
            if use_bands and not self.compact and g.total_px >= 2 ** 32:
                raise ValueError("a batch with dense outputs must hold fewer than 2**32 pixels (split it, or use compact=True)")
            if use_bands:
                from .morphology import pass_radius
                halo = sum(pass_radius(t) for t, _ in passes)
                d_bands, d_band_off, n_bands, left, bands_h, band_off_h = batch.band_lists(halo)
                run_cap = max(g.total_words // 3, 1 << 16)
                runs = ws.get("runs", run_cap, torch.int64, dev)          # maze_run_t, 8 bytes each
                run_stats = ws.get("run_stats", run_cap, torch.int64, dev)
                run_pix = ws.get("run_pix", run_cap, torch.int32, dev)
                band_out = ws.get("band_out", 4 * max(n_bands, 1), torch.int32, dev)
                band_counters = ws.get("band_counters", 8, torch.int32, dev)
                big_list = ws.get("big_list", 3 * n, torch.int32, dev)
                band_done = ws.get("band_done", n, torch.int32, dev)
                # frames (>= HUGE_PX pixels) are labelled by the global-memory kernels: {vignette, its number of bands}
                from ._lib import HUGE_PX
                nb_of = np.diff(band_off_h)
                huge = np.nonzero((g.npx >= HUGE_PX) & (nb_of > 0))[0]
                huge_pairs = np.ascontiguousarray(np.stack([huge, nb_of[huge]], axis=1).astype(np.int32)) if len(huge) else None
                gl_scratch = ws.get("gl_scratch", 2 * run_cap + n_bands + 16, torch.int32, dev) if len(huge) else None
            else:
                d_list, class_off, left = batch.fused_lists()
            # a step that is a handful of vignettes (a frame, BASELINE.json configs[3]) is some twenty small launches whose
            # issue time exceeds the GPU time: it is captured once per argument block and replayed as a CUDA graph
            use_graph = self.graphs and use_bands and len(left) == 0 and n <= 16
            # rotating pinned readback slots (one more than lanes): a slot is reused only after its batch was finalised.
            # (Graph steps replay their arguments: one slot per lane, rewritten when the lane comes round again -- by then
            # the lane's previous result is past its validity anyway.)
            if use_graph:
                slots, ri = self._readback_graph, getattr(ws, "index", 0)
                nslot = self.n_lanes
            else:
                slots, nslot = self._readback, self.n_lanes + 1
                ri = self._readback_i % nslot
                self._readback_i += 1
            while len(slots) < nslot:
                slots.append(None)
            slot = slots[ri]
            if slot is None or slot.numel() < 3 * n + 2:
                slot = torch.empty(3 * n + 2 + 256, dtype=torch.int32, pin_memory=True)
                slots[ri] = slot
            host = slot[:3 * n + 2]
            # the whole step is ONE call into the library (maze_stage_step): band pipeline (or fused kernel) on the lane
            # stream, the oversize vignettes through the per-operator chain on the side stream, offsets, feature rows, readback
            a = StepArgs()
            a.vig = batch.d_vig.data_ptr()
            a.image, a.intensity = d_src.data_ptr(), d_image.data_ptr()
            a.bits, a.mask, a.labels = bits.data_ptr(), mask.data_ptr(), labels.data_ptr()
            a.counts, a.lab_off, a.stage_counter = counts.data_ptr(), lab_off.data_ptr(), staging[3].data_ptr()
            a.acc_stage, a.hi_stage, a.ext_stage = staging[0].data_ptr(), staging[1].data_ptr(), staging[2].data_ptr()
            a.table, a.counts_host = table.data_ptr(), host.data_ptr()
            if use_bands:
                a.bands, a.band_off, a.n_bands, a.halo = d_bands.data_ptr(), d_band_off.data_ptr(), n_bands, halo
                a.runs, a.run_stats, a.run_pix = runs.data_ptr(), run_stats.data_ptr(), run_pix.data_ptr()
                a.band_out, a.band_counters, a.big_list = band_out.data_ptr(), band_counters.data_ptr(), big_list.data_ptr()
                a.run_cap, a.total_px = run_cap, g.total_px
                a.band_done = band_done.data_ptr()
                pp = self.postprocess
                if pp is not None:
                    a.clear_border, a.min_area = int(bool(pp.clear_border)), int(pp.min_area)
                if huge_pairs is not None:
                    a.huge_host, a.n_huge, a.huge_px = huge_pairs.ctypes.data, len(huge_pairs), HUGE_PX
                    a.gl_scratch = gl_scratch.data_ptr()
                a.step_flags = STEP_COMPACT if self.compact else 0
            else:
                a.img_list = d_list.data_ptr()
                for c in range(len(class_off)):
                    a.class_off[c] = int(class_off[c])
            for k, (t, inv) in enumerate(passes):
                a.pass_t[k], a.pass_invert[k] = int(t), int(inv)
            a.n_img, a.t_int, a.n_pass, a.stage_cap = n, int(t_int), len(passes), cap
            a.flags = RP_HIGH_ORDER if self.high_order else 0
            single = self.postprocess is None  # threshold branch: the whole mask is one region (ImageProperties)
            if single:
                a.flags |= BAND_SINGLE_REGION
            a.left_n = len(left)
            keep = None
            if len(left):
                sub, _, idx = self._sub(batch, left)
                tiles_full = batch.tiles_of(left)
                ar = ws.arena
                keep = (ar.take(max(g.total_words, 1), torch.int32), ar.take(2 * len(left), torch.int32),
                        ar.take(g.total_px, torch.int32), ar.take(sub.g.n_tiles + 1, torch.int32),
                        ar.take(len(left) + 1, torch.int32), ar.take(cap * NACC, torch.int64), ar.take(cap * NEXT, torch.int32))
                a.left_vig, a.left_tiles, a.left_idx = sub.d_vig.data_ptr(), sub.d_tiles.data_ptr(), idx.data_ptr()
                a.left_tiles_full = tiles_full.data_ptr()
                a.left_n_tiles, a.left_n_tiles_full = sub.g.n_tiles, tiles_full.numel() // 8
                (a.scratch_plane, a.scratch_flags, a.scratch_parent, a.scratch_tile_scan, a.scratch_lab_off, a.scratch_acc,
                 a.scratch_ext) = (t.data_ptr() for t in keep)
            entry = None
            if use_graph:
                saved_hh, a.huge_host = a.huge_host, 0
                key = (bytes(a), huge_pairs.tobytes() if huge_pairs is not None else b"")
                a.huge_host = saved_hh
                graphs = ws.__dict__.setdefault("graphs", {})
                entry = graphs.get(key)
                if entry is None:
                    if len(graphs) >= 16:  # (a stream of ever new batches: the oldest argument block goes)
                        old = graphs.pop(next(iter(graphs)))
                        if old[1].value:
                            lib().maze_graph_destroy(old[1])
                    entry = graphs[key] = [0, ctypes.c_void_p(0), ctypes.c_int(0)]  # times seen, exec handle, launches inside
            ws.plan = None
            if use_graph and entry is not None:
                ws.plan = (ck, batch, (bits, mask, labels, counts, lab_off, n_labels, acc_base, cap, staging, table,
                                       use_bands, bands_h, band_off_h, runs, band_out, left, host, a, keep, huge_pairs, entry,
                                       single, use_graph))
        if entry is not None and entry[0] >= 1 and entry[2].value >= 0:
            # (the first step with these arguments runs plainly -- it also creates what the library creates lazily --
            # the second one is captured, the following ones are one graph launch)
            check(lib().maze_stage_step_graph(ctypes.byref(a), main.cuda_stream, side.cuda_stream, ctypes.byref(entry[1]),
                                              ctypes.byref(entry[2])), "maze_stage_step_graph")
        else:
            check(lib().maze_stage_step(ctypes.byref(a), main.cuda_stream, side.cuda_stream), "maze_stage_step")
        if entry is not None:
            entry[0] += 1
        side_done = None
        if len(left):
            side_done = torch.cuda.Event()
            side_done.record(side)
            ws.side_done = side_done
        done = torch.cuda.Event()
        done.record(main)
        left_set = set(int(i) for i in left)

        def pending(res):
            done.synchronize()
            if side_done is not None:
                side_done.synchronize()
            h = host.numpy()
            total = int(h[3 * n])
            res.n_runs = int(h[3 * n + 1]) if use_bands else 0
            bad = np.nonzero((h[n:2 * n] != 0) | ((h[2 * n:3 * n] < 0) & (h[:n] > 0)))[0]
            bad = [int(i) for i in bad if int(i) not in left_set]
            res.dense_only = sorted(left_set | set(bad))  # vignettes without a run list (per-operator kernels)
            ppf = self.postprocess
            if bad and ppf is None and total <= cap and len(bad) <= 256:
                # threshold branch: the few vignettes whose run tables overflowed (speckle) are thresholded again with
                # the per-operator kernels, in place; every vignette keeps exactly one row
                with torch.cuda.stream(main):
                    res.redone = True
                    sub, word_idx, idx = self._sub(batch, bad)
                    sub.arena = batch.arena
                    sbits, _ = sub.threshold_pack(d_src, t_int)
                    bits[word_idx] = sbits[word_idx]
                    sub.unpack_mask(sbits, out=mask)
                    n_labels[idx.long()] = 1
                    batch.count_scan(n_labels, out=lab_off)
                    batch.props_finish_staged(staging, acc_base, lab_off, cap, True, self.high_order, table)
                    batch.regionprops(lab_off, cap, labels=None, bits=bits, image=d_image, high_order=self.high_order,
                                      table=table, acc_base=acc_base, tiles=batch.tiles_of(sorted(bad)))
                    total = n
                    res._table = table[:total]
                    res._sync_main = True
                    res.ready = torch.cuda.Event()
                    res.ready.record(main)
            elif (bad or total > cap) and ppf is None:
                # threshold branch: the whole batch through the per-operator kernels
                with torch.cuda.stream(main):
                    r2 = self.run_device_threshold_generic(batch, d_src, d_image, t_int)
                    res.redone = True
                    res.dense_only = list(range(n))
                    res.bits, res.mask, res.lab_off, res.keep = r2.bits, r2.mask, r2.lab_off, r2.keep
                    res._table, total = r2._table, r2._n_obj
                    res._sync_main = True
                    res.ready = torch.cuda.Event()
                    res.ready.record(main)
            elif (bad or total > cap) and (ppf.clear_border or ppf.min_area > 0):
                # the per-operator redo below knows no label filters: the whole batch takes the filter path instead
                with torch.cuda.stream(main):
                    r2 = self._run_filter_path(batch, d_src, d_image, t_int, passes)
                    res.redone = True
                    res.dense_only = list(range(n))
                    res.bits, res.labels, res.mask, res.lab_off = r2.bits, r2.labels, r2.mask, r2.lab_off
                    res._table, total = r2._table, r2._n_obj
                    res._sync_main = True
                    res.ready = torch.cuda.Event()
                    res.ready.record(main)
            elif bad or total > cap:
                # some vignettes overflowed the fused kernel's tables (more runs than slots, or the staging rows
                # ran out): the per-operator kernels redo just those, then offsets and features are re-derived
                with torch.cuda.stream(main):
                    res.redone = True
                    total2 = cap + 1
                    if total <= cap and len(bad) <= 256:
                        self._redo_generic(batch, bad, d_src, t_int, bits, mask, labels, n_labels)
                        batch.count_scan(n_labels, out=lab_off)
                        total2 = int(lab_off[n].item())
                    if total2 <= cap:
                        batch.props_finish_staged(staging, acc_base, lab_off, cap, True, self.high_order, table)
                        todo = sorted(left_set | set(bad))
                        batch.regionprops(lab_off, cap, labels=labels, bits=bits, image=d_image,
                                          high_order=self.high_order, runs=True, table=table, acc_base=acc_base,
                                          tiles=batch.tiles_of(todo))
                        total = total2
                        res._table = table[:total]
                    else:  # the whole batch through the per-operator kernels
                        res.dense_only = list(range(n))
                        b2, l2, off2, m2 = self._front_generic(batch, d_src, t_int, labels=labels, mask=mask)
                        total = int(off2[-1].item())
                        res.bits, res.lab_off = b2, off2
                        res._table = batch.regionprops(off2, total, labels=l2, bits=b2, image=d_image,
                                                       high_order=self.high_order, runs=True)
                    res._sync_main = True
                    res.ready = torch.cuda.Event()
                    res.ready.record(main)  # the redo ran on the lane stream
            else:
                res._table = table[:total]
            res._n_obj = total

        res = DeviceResult(batch, bits, None if single else labels, lab_off, None, None, mask=mask, pending=pending)
        res.single_region = single
        res.ready = done
        if use_bands:
            res.runs, res.band_out, res.bands_host, res.band_off_host = runs, band_out, bands_h, band_off_h
            res.compact = self.compact
        return res

    def reserve(self, geometries, device=None):
        """Size both device workspaces for the largest of the given batch geometries, so that the steady
        state makes no allocator call (cudaMalloc inside a step costs milliseconds)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        px = max(g.total_px for g in geometries)
        words = max(max(g.total_words, 1) for g in geometries)
        n = max(g.n_img for g in geometries)
        cap = max(_stage_cap(g) for g in geometries)
        for ws in self._ws_ring + [self._ws]:
            for key, size, dt in (("bits", words, torch.int32), ("mask", px, torch.uint8), ("labels", px, torch.int32),
                                  ("counts", 3 * n, torch.int32), ("lab_off", n + 1, torch.int32),
                                  ("acc", cap * NACC, torch.int64), ("hi", cap * 8, torch.float64),
                                  ("ext", cap * NEXT, torch.int32), ("counter", 1, torch.int32),
                                  ("table", cap * NFEAT, torch.float64), ("runs", max(words // 3, 1 << 16), torch.int64),
                                  ("run_stats", max(words // 3, 1 << 16), torch.int64), ("run_pix", max(words // 3, 1 << 16), torch.int32),
                                  ("band_out", 4 * max(2 * n, words // 1024 + n), torch.int32),
                                  ("band_counters", 8, torch.int32), ("big_list", 3 * n, torch.int32), ("band_done", n, torch.int32)):
                ws.get(key, size, dt, dev)
            if getattr(ws, "arena", None) is None or ws.arena.device != dev:
                ws.arena = Arena(dev)
            # (the scratch arena of the per-operator path -- oversize vignettes, redo of flagged ones -- gets a token
            # reservation only: with the band pipeline that path is the exception, and batch-sized scratch for every
            # lane costs several GB per lane; the arena grows on first use)
            ws.arena.reserve(8 << 20)
            if getattr(ws, "lane", None) is None or ws.lane.device != dev:
                ws.lane = torch.cuda.Stream(device=dev)
                ws.side = torch.cuda.Stream(device=dev)

    def prepare(self, batch: DeviceBatch):
        """Build the per-batch launch plan (size classes, descriptors of the vignettes that need the
        per-operator path) ahead of run_device; part of making a batch resident."""
        if self.postprocess is not None and self.fused and self._passes() is not None:
            if self.pipeline == "bands":
                from .morphology import pass_radius
                left = batch.band_lists(sum(pass_radius(t) for t, _ in self._passes()))[3]
            else:
                left = batch.fused_lists()[2]
            if len(left):
                self._sub(batch, left)
                batch.tiles_of(left)
        return batch

    def _sub(self, batch, indices):
        key = tuple(int(i) for i in indices)
        cache = batch.__dict__.setdefault("_sub_cache", {})
        if key not in cache:
            g = batch.g
            sub = DeviceBatch(g.subset(indices), batch.device)
            word_idx = np.concatenate([np.arange(g.word_off[i], g.word_off[i] + g.nwords[i]) for i in indices])
            cache[key] = (sub, torch.from_numpy(word_idx).to(batch.device),
                          torch.as_tensor(np.asarray(indices, np.int32), device=batch.device))
        return cache[key]

    def _redo_generic(self, batch, indices, d_src, t_int, bits, mask, labels, n_labels):
        """Run the per-operator kernels on a few vignettes of the batch, in place in the batch buffers."""
        sub, word_idx, idx = self._sub(batch, indices)
        sub.arena = batch.arena
        passes = self._passes()
        if passes is not None:
            sbits, slab_off = sub.front_chain(d_src, t_int, passes, labels, mask)
        else:
            sbits, _, slab_off, _ = self._front_generic(sub, d_src, t_int, labels=labels, mask=mask)
        bits[word_idx] = sbits[word_idx]
        n_labels[idx.long()] = slab_off[1:] - slab_off[:-1]

    # ---- host entry: numpy in, numpy out --------------------------------------------------------------
    def _halo(self):
        """Halo of the band plan (sum of the pass radii) when the asynchronous band pipeline will take the batch."""
        pp = self.postprocess
        if pp is None:
            return 0 if (self.fused and self.pipeline == "bands") else None
        if not self.fused or self.pipeline != "bands" or self._passes() is None:
            return None
        if pp.merge_segments_distance > 0:
            return None
        from .morphology import pass_radius
        return sum(pass_radius(t) for t, _ in self._passes())

    def _prepack(self, images, foreground_pred, pool):
        """Geometry of the batch and the START of its packing into the pinned staging buffer (copy threads run in the
        background); _enqueue waits for them right before the upload."""
        if self.threshold is None and foreground_pred is None:
            raise ValueError("postprocess-only stage needs foreground_pred")
        geom = BatchGeometry.from_images(images)
        if geom.n_img == 0:
            return (geom, None, None, None)
        h_img = pool.get("img", geom.total_px, torch.uint8)
        wait = geom.pack_host_start(images, out=h_img.numpy())
        h_pred = None
        if self.threshold is None:
            h_pred = pool.get("pred", geom.total_px, torch.uint8)
            geom.pack_host([np.asarray(p, dtype=bool).view(np.uint8) for p in foreground_pred], out=h_pred.numpy())
        return (geom, h_img, h_pred, wait)

    def _enqueue(self, images, foreground_pred, pool, want_mask, want_labels, pre=None):
        """Upload one (pre)packed batch, launch its kernels and start the download of the per-pixel outputs on the
        copy stream; returns what _complete needs.  Nothing here waits for the GPU."""
        if pre is None:
            pre = self._prepack(images, foreground_pred, pool)
        geom, h_img, h_pred, wait = pre
        if geom.n_img == 0:
            return (geom, None, None, None, None, None, None)
        dev = self.device
        batch = DeviceBatch(geom, dev)
        # descriptors and launch plan go up BEFORE the image, through one pinned buffer and one asynchronous copy
        batch.upload_descriptors(pool, self._halo())
        self.prepare(batch)
        main = torch.cuda.current_stream()
        if self._copy_stream is None or self._copy_stream.device != batch.device:
            self._copy_stream = torch.cuda.Stream(device=batch.device)
        wait()  # the staging buffer is complete
        d_image = pool.dev("img", geom.total_px, torch.uint8, batch.device)
        d_image.copy_(h_img, non_blocking=True)
        d_pred = None
        if h_pred is not None:
            d_pred = pool.dev("pred", geom.total_px, torch.uint8, batch.device)
            d_pred.copy_(h_pred, non_blocking=True)
        res = self.run_device(batch, d_image, d_pred)
        if getattr(res, "deferred_pixels", False):
            res.finalize()  # (merge_labels rewrites the label image in finalize)
        # the per-pixel outputs do not depend on the object counts: download them on the copy stream while
        # the next batch is packed, uploaded and computed
        cs = self._copy_stream
        if res.ready is not None:
            cs.wait_event(res.ready)
        else:
            cs.wait_stream(main)
        h_mask = h_lab = None
        if getattr(res, "compact", False):
            want_mask = want_labels = False  # the run list goes down in _complete, once the number of runs is known
        with torch.cuda.stream(cs):
            if want_mask:
                h_mask = pool.get("mask", geom.total_px, torch.uint8)
                h_mask.copy_(res.mask, non_blocking=True)
            if want_labels and res.labels is not None:
                h_lab = pool.get("labels", geom.total_px, torch.int32)
                h_lab.copy_(res.labels, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(cs)
        if res._sync_main:  # produced in the shared workspace: the next batch on that path must wait for this download
            self._shared_busy = copied
        return (geom, batch, res, h_mask, h_lab, pool, (d_image, d_pred, copied))

    def _complete(self, inflight) -> StageResult:
        geom, batch, res, h_mask, h_lab, pool, _keepalive = inflight
        if batch is None:
            return StageResult(geom, np.zeros(0, np.uint8), np.zeros(0, np.int32), np.zeros(1, np.int32),
                               np.zeros((0, NFEAT)))
        cs = self._copy_stream
        if self._small_copy_stream is None or self._small_copy_stream.device != batch.device:
            self._small_copy_stream = torch.cuda.Stream(device=batch.device)
        ss = self._small_copy_stream  # table / offsets must not queue behind the NEXT batch's big downloads on cs
        main = torch.cuda.current_stream()
        copied = _keepalive[2]
        res.finalize()  # waits for the counts; redoes the vignettes that overflowed the fused kernel
        if res._sync_main:  # the table (and, after a redo, the per-pixel outputs) came from kernels no host sync covered
            if res.ready is not None:
                ss.wait_event(res.ready)
            else:
                ss.wait_stream(main)
        compact = getattr(res, "compact", False) and res.runs is not None and len(res.dense_only) <= 64
        merged = None
        if compact and getattr(res, "n_merge", None) is not None:
            # merge_labels rewrote the label images of the vignettes in which labels merged: the run list of the band
            # step still gives their MASKS (merge_labels.py never touches foreground_pred) and everything of the other
            # vignettes; the merged label images come down dense -- unless that is most of the batch anyway
            with torch.cuda.stream(ss):
                merged = np.nonzero(res.n_merge.cpu().numpy() > 0)[0]
            if int(geom.npx[merged].sum()) * 2 > geom.total_px:
                compact, merged = False, None
        if getattr(res, "compact", False) and not compact:
            with torch.cuda.stream(ss):  # (rare) most of the batch has no run list: dense download after all
                h_mask = pool.get("mask", geom.total_px, torch.uint8)
                h_mask.copy_(res.mask, non_blocking=True)
                h_lab = pool.get("labels", geom.total_px, torch.int32)
                h_lab.copy_(res.labels, non_blocking=True)
        h_runs = h_bo = None
        dense = {}
        if compact:
            from .device import BAND_OUT_DTYPE, RUN_DTYPE
            n_runs = min(int(res.n_runs), res.runs.numel())
            n_bands = len(res.bands_host)
            with torch.cuda.stream(ss):
                h_runs = pool.get("runs", max(n_runs, 1), torch.int64)[:n_runs]
                h_runs.copy_(res.runs[:n_runs], non_blocking=True)
                h_bo = pool.get("band_out", 4 * max(n_bands, 1), torch.int32)[:4 * n_bands]
                h_bo.copy_(res.band_out[:4 * n_bands], non_blocking=True)
                if merged is not None and len(merged):
                    sub_off = np.concatenate([[0], np.cumsum(geom.npx[merged])]).astype(np.int64)
                    h_sub = pool.get("labels_merged", max(int(sub_off[-1]), 1), torch.int32)
                    for k, i in enumerate(merged):
                        o, npx = int(geom.pix_off[i]), int(geom.npx[i])
                        h_sub[int(sub_off[k]):int(sub_off[k]) + npx].copy_(res.labels[o:o + npx], non_blocking=True)
                    h_sub_np = h_sub.numpy()
                    for k, i in enumerate(merged):
                        dense[int(i)] = (None, h_sub_np[int(sub_off[k]):int(sub_off[k + 1])].reshape(int(geom.h[i]),
                                                                                                    int(geom.w[i])))
                for i in res.dense_only:  # vignettes of the per-operator kernels: their dense arrays
                    o, npx = int(geom.pix_off[i]), int(geom.npx[i])
                    shp = (int(geom.h[i]), int(geom.w[i]))
                    dense[i] = (res.mask[o:o + npx].cpu().numpy().reshape(shp),
                                None if res.labels is None else res.labels[o:o + npx].cpu().numpy().reshape(shp))
        with torch.cuda.stream(ss):
            if res.redone and not compact:  # rare: per-pixel outputs were rewritten by the per-operator path
                ss.wait_event(copied)
                if h_mask is not None:
                    h_mask.copy_(res.mask, non_blocking=True)
                if h_lab is not None:
                    h_lab.copy_(res.labels, non_blocking=True)
            table = res.table
            h_tab = pool.get("table", max(table.numel(), 1), torch.float64)[:table.numel()]
            h_tab.copy_(table.reshape(-1), non_blocking=True)
            h_off = pool.get("lab_off", geom.n_img + 1, torch.int32)
            h_off.copy_(res.lab_off, non_blocking=True)
            h_shape = None
            if self.shape_features:
                # (after merge_labels the labels leave the runs of the bit plane: per-pixel plane build)
                d_shape = batch.label_shape(table, labels=res.labels, bits=res.bits, runs=res.merge_status is None)
                h_shape = pool.get("shape", max(d_shape.numel(), 1), torch.float64)[:d_shape.numel()]
                h_shape.copy_(d_shape.reshape(-1), non_blocking=True)
            keep = None if res.keep is None or not torch.is_tensor(res.keep) else res.keep.cpu().numpy()
            status = None if res.merge_status is None else res.merge_status.cpu().numpy()
        ss.synchronize()
        copied.synchronize()
        failed = None
        if status is not None and (status == MAZE_ERR_TYPEERROR).any():
            if self.merge_errors == "raise":
                # the reference aborts the run here (merge_labels.py:19-20 via pipeline_runner.py:40-43)
                raise TypeError("'NoneType' object is not iterable")
            failed = np.nonzero(status == MAZE_ERR_TYPEERROR)[0]
        if compact:
            from .device import BAND_OUT_DTYPE, RUN_DTYPE
            out = StageResult.from_runs(geom, h_runs.numpy().view(RUN_DTYPE), h_bo.numpy().view(BAND_OUT_DTYPE),
                                        res.band_off_host, np.ascontiguousarray(
                                            _rpb_of(res.bands_host, res.band_off_host, geom.n_img)), dense,
                                        h_off.numpy(), h_tab.numpy().reshape(-1, NFEAT))
        else:
            out = StageResult(geom, None if h_mask is None else h_mask.numpy(), None if h_lab is None else h_lab.numpy(),
                              h_off.numpy(), h_tab.numpy().reshape(-1, NFEAT), keep=keep)
        if res.single_region and keep is None:
            # Filter(obj[mask].any()), loki/pipeline.py:651: a vignette is kept when its (single) region has pixels
            t_np, o_np = h_tab.numpy().reshape(-1, NFEAT), h_off.numpy()
            keep = t_np[o_np[:-1], 1] > 0 if len(t_np) else np.zeros(geom.n_img, bool)
            out.keep = keep
            out._no_labels = True
        out.merge_failed = failed
        if self.shape_features:
            out.shape_table = h_shape.numpy().reshape(-1, NSHAPE)
        return out

    def __call__(self, images: Sequence[np.ndarray], foreground_pred: Optional[Sequence[np.ndarray]] = None,
                 want_mask=True, want_labels=True) -> StageResult:
        """One batch: list of uint8 (h, w) vignettes in, masks / label images / object table out (numpy views of
        pinned buffers that stay valid until the next call)."""
        dev = self.device
        with torch.cuda.device(dev if dev is not None else torch.cuda.current_device()):
            return self._complete(self._enqueue(images, foreground_pred, self._pool, want_mask, want_labels))

    def map(self, batches, want_mask=True, want_labels=True):
        """Streaming form: ``for res in stage.map(iterable_of_image_lists)``.  Packing and upload of batch i+1
        overlap the kernels and the download of batch i (three rotating pinned buffer sets, rotating device
        workspaces, a separate copy stream).  A yielded result stays valid until the next-but-one is yielded.
        Items may be image lists or (images, foreground_pred) pairs."""
        dev = self.device
        with torch.cuda.device(dev if dev is not None else torch.cuda.current_device()):
            pending = None
            # batch i+1 may only be enqueued before batch i is complete when the two use different device buffers:
            # the asynchronous band / fused path rotates its workspaces (>= 2 lanes); the paths with label filters,
            # merge_labels, radii beyond the bit-plane kernels or threshold-only work in ONE shared workspace
            pp = self.postprocess
            if pp is None:
                overlap = self.fused and self.n_lanes >= 2 and self.pipeline == "bands"
            else:
                overlap = (self.fused and self.n_lanes >= 2 and self._passes() is not None
                           and pp.merge_segments_distance <= 0
                           and (self.pipeline == "bands" or not (pp.clear_border or pp.min_area > 0)))
            def split(item):
                return item if isinstance(item, tuple) else (item, None)

            # Batches are PACKED two ahead: the copy threads of batch i+1 and i+2 fill their pinned staging buffers in the
            # background while batch i computes, batch i-1 is completed and the consumer works on what was yielded.
            # Staging buffers rotate over four sets: the set of batch i+2 was last used by batch i-2, which is complete.
            if self._img_ring is None:
                self._img_ring = [_PinnedPool() for _ in range(4)]
            it = iter(batches)
            ahead = []  # prepacked batches, oldest first
            n_pre = 0

            def prefetch():
                nonlocal n_pre
                while len(ahead) < 2:
                    item = next(it, None)
                    if item is None:
                        return
                    ahead.append(self._prepack(*split(item), self._img_ring[n_pre % 4]))
                    n_pre += 1

            i = 0
            try:
                prefetch()
                while ahead:
                    if not overlap and pending is not None:
                        yield self._complete(pending)
                        pending = None
                    pre = ahead.pop(0)
                    inflight = self._enqueue(None, None, self._map_pools[i % 3], want_mask, want_labels, pre=pre)
                    prefetch()
                    if pending is not None:
                        yield self._complete(pending)
                    pending = inflight
                    i += 1
                if pending is not None:
                    yield self._complete(pending)
            finally:
                for pre in ahead:  # never leave copy threads behind
                    if pre[3] is not None:
                        pre[3]()


def _merge_candidates(tab, off, hs, ws, max_distance):
    """Vignettes merge_labels may change (see LokiSegmentationStage._run_merge_on_bands): those with two or more live
    labels in which some label's bounding box comes within max_distance of the smallest label's in both directions, or
    in which a label's distance window is clipped by the frame on all four sides.  tab: columns label, area, bbox of the
    object table; off: first row of every vignette."""
    import math
    n = len(off) - 1
    if len(tab) == 0:
        return np.zeros(0, np.int64)
    img = np.repeat(np.arange(n), np.diff(off))
    live = tab[:, 1] > 0
    rows = np.nonzero(live)[0]
    if len(rows) == 0:
        return np.zeros(0, np.int64)
    iv = img[rows]
    cnt = np.bincount(iv, minlength=n)
    first = np.full(n, -1, np.int64)
    first[iv[::-1]] = rows[::-1]                      # first live row of every vignette = its smallest label l0
    r0, c0, r1, c1 = (tab[rows, 2 + k] for k in range(4))           # r1 / c1 exclusive
    f = first[iv]
    R0, C0, R1, C1 = tab[f, 2], tab[f, 3], tab[f, 4], tab[f, 5]
    dy = np.maximum(np.maximum(r0 - (R1 - 1), R0 - (r1 - 1)), 0)    # rows between the nearest pixel rows
    dx = np.maximum(np.maximum(c0 - (C1 - 1), C0 - (c1 - 1)), 0)
    near = (np.maximum(dy, dx) <= max_distance) & (rows != f)
    pad = math.ceil(max_distance) + 1
    H, W = np.asarray(hs)[iv], np.asarray(ws)[iv]
    clipped = (r0 - pad < 0) & (r1 - 1 + pad > H - 1) & (c0 - pad < 0) & (c1 - 1 + pad > W - 1)
    flag = np.zeros(n, bool)
    np.logical_or.at(flag, iv, near | clipped)
    return np.nonzero(flag & (cnt >= 2))[0].astype(np.int64)


def _stage_cap(g):
    """Rows of the staging arrays / object table of one batch: 16 objects per vignette on average plus, for frames
    (BASELINE.json configs[3]: thousands of labels in one image), one per 1024 pixels."""
    from ._lib import HUGE_PX
    big = g.npx[g.npx >= HUGE_PX]
    return 16 * g.n_img + 1024 + int((big // 1024).sum())


def _rpb_of(bands, band_off, n_img):
    """Rows per band of every vignette (1 for vignettes without bands)."""
    rpb = np.ones(n_img, np.int32)
    has = band_off[1:] > band_off[:-1]
    rpb[has] = bands["rpb"][band_off[:-1][has]]
    return rpb


def stream_objects(stage: "LokiSegmentationStage", objects, batch_size: int = 2048, image_key: str = "image",
                   meta_key: str = "meta", padding: int = 75, min_intensity=None, object_id_fmt=None,
                   keep_arrays: bool = True, rois: bool = False, apply_mask: bool = False, background_color=0,
                   keep_background: bool = True):
    """Adapter for a morphocut-style object stream (the place of the `Call` chain of
    maze_ipp/loki/pipeline.py:396-459 and the FindRegions / recalc_metadata / CalculateZooProcessFeatures tail,
    :589-625): consumes an iterable of dict-like stream objects that carry a uint8 vignette or frame under
    `image_key`, buffers `batch_size` of them, runs the stage once per buffer (pipelined over buffers) and yields,
    in input order, one dict per input object with the keys of the input plus `mask`, `labels` (`keep_arrays`) and
    `objects` (a list of metadata dicts, one per segmented object, see regions.objects_of).  In the threshold branch
    (loki/pipeline.py:648-656) every vignette that survives the empty-mask filter yields ONE object, the whole mask.
    ``rois=True`` adds ``rois``: per object the ExtractROI crop (``apply_mask`` / ``background_color`` /
    ``keep_background`` as in the schema) and the object's mask inside it."""
    from .regions import extract_roi, find_regions, recalc_metadata, zooprocess_features

    def buffers():
        buf = []
        for obj in objects:
            buf.append(obj)
            if len(buf) == batch_size:
                yield buf
                buf = []
        if buf:
            yield buf

    pending = []

    def images():
        for buf in buffers():
            pending.append(buf)
            imgs = []
            for o in buf:
                im = np.asarray(o[image_key])
                if im.dtype != np.uint8 or im.ndim != 2:  # ImageReader(picture_fn, "L") delivers uint8 (pipeline.py:919-921)
                    raise TypeError(f"stream_objects needs 2-D uint8 images, got {im.dtype} with shape {im.shape}")
                imgs.append(np.ascontiguousarray(im))
            yield imgs

    threshold_only = stage.postprocess is None
    for res in stage.map(images()):
        buf = pending.pop(0)
        for i, obj in enumerate(buf):
            if res.keep is not None and not res.keep[i]:
                continue  # Filter(obj[mask].any()), loki/pipeline.py:651: vignettes with an empty mask are dropped
            out = dict(obj)
            meta = obj.get(meta_key, {}) if hasattr(obj, "get") else {}
            if keep_arrays:
                out["mask"] = res.mask(i).copy()
                lab = res.labels(i)
                out["labels"] = None if lab is None else lab.copy()
            # threshold branch: ImageProperties(mask, image) is ONE region over the whole frame, no padding (:653)
            regs = list(find_regions(res, i, 0 if threshold_only else padding, None if threshold_only else min_intensity,
                                     obj[image_key]))
            out["objects"] = [zooprocess_features(r, recalc_metadata(r, meta, object_id_fmt)) for r in regs]
            if rois:
                # ExtractROI (loki/pipeline.py:596-602, options of config_schema.py:90-107): the vignette and the mask
                # of every object, what the EcoTaxa writer stores (`image`, and `mask` with store_mask)
                out["rois"] = [(extract_roi(obj[image_key], r, 1 if apply_mask else 0, background_color, keep_background),
                                r.image) for r in regs]
            yield out


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` of `world` (images are independent)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_object_tables(table: np.ndarray, image_index0: int, group=None):
    """Concatenate the per-rank object tables on rank 0 in image order (no collective on the hot path:
    this is the host-side gather of the results).  Column MAZE_F_IMAGE is rebased to global indices."""
    import torch.distributed as dist

    local = table.copy()
    if local.size:
        local[:, 57] += image_index0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    parts = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if parts is None:
        return None
    return np.concatenate([p.reshape(-1, NFEAT) for p in parts], axis=0)
