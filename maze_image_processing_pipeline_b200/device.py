"""Packed-batch geometry and the device-side operators (thin wrappers over the C-ABI).

PyTorch is used for device memory, streams and pinned host buffers only; every computation is a
kernel of ``libmaze_b200.so``.  Layout (see ``include/maze_b200.h``): all vignettes of a batch are
concatenated in flat per-pixel arrays (numpy C order per vignette, vignette starts aligned to 16
elements); binary images are bit planes with rows padded to 32-bit words; the batch is cut into
256-word tiles, one CTA each.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import (BAND_PLANE_WORDS, HUGE_PX, FUSED_CAPS, FUSED_NO_PROPS, MAX_DISK_RADIUS, NACC, NEXT, NFEAT, RP_HIGH_ORDER, RP_RUNS,
                   TILE_WORDS, check, lib)

VIG_DTYPE = np.dtype([("pix_off", "<i8"), ("word_off", "<i8"), ("h", "<i4"), ("w", "<i4"),
                      ("wpr", "<i4"), ("tile0", "<i4")])
TILE_DTYPE = np.dtype([("img", "<i4"), ("word0", "<i4")])
BAND_TARGET_CTAS = 444  # 148 SMs x 3
BAND_DTYPE = np.dtype([("img", "<i4"), ("y0", "<i4"), ("y1", "<i4"), ("rpb", "<i4")])        # maze_band_t
BAND_OUT_DTYPE = np.dtype([("base", "<i4"), ("n_runs", "<i4"), ("zflags", "<i4"), ("reserved", "<i4")])  # maze_band_out_t
RUN_DTYPE = np.dtype([("y", "<u2"), ("x0", "<u2"), ("x1", "<u2"), ("label", "<u2")])            # maze_run_t
assert VIG_DTYPE.itemsize == 32 and TILE_DTYPE.itemsize == 8 and BAND_DTYPE.itemsize == 16 and RUN_DTYPE.itemsize == 8

_DISK_T_LIMIT = (MAX_DISK_RADIUS + 1) ** 2
WIDE_MAX_R = 254                 # maze_morph_pass_wide: squared-distance thresholds up to 254^2
WIDE_MIN_R = int(os.environ.get("MAZE_WIDE_MIN_R", "22"))  # from this radius on the separable pass pair is the faster one
# (2048^2 frame, per pass: bit-plane kernel 0.027 / 0.045 / 0.10 ms at r = 8 / 16 / 32; separable pair 0.06 / 0.065 / 0.077 /
# 0.11 ms at r = 8 / 16 / 32 / 64 -- tools/wide_probe.py)
_INT_MAX = 2 ** 31 - 1


class BatchGeometry:
    """Host-side description of a packed batch (numpy only, no device access)."""

    def __init__(self, heights, widths, _pix_off=None, _word_off=None, _total_px=None, _total_words=None):
        hs = np.asarray(heights, dtype=np.int64).ravel()
        ws = np.asarray(widths, dtype=np.int64).ravel()
        if hs.shape != ws.shape:
            raise ValueError("heights and widths differ in length")
        if hs.size and (hs.min() < 1 or ws.min() < 1):
            raise ValueError("empty vignettes cannot be packed")
        if hs.size and (hs * ws).max() >= 2 ** 31:
            raise ValueError("a vignette must have fewer than 2**31 pixels")
        self.n_img = int(hs.size)
        self.h = hs
        self.w = ws
        self.npx = hs * ws
        wpr = (ws + 31) // 32
        nwords = hs * wpr
        padded = (self.npx + 15) // 16 * 16
        self.nwords = nwords
        self.pix_off = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
        self.word_off = np.concatenate([[0], np.cumsum(nwords)]).astype(np.int64)
        if _pix_off is not None:  # a subset that addresses its parent's buffers
            self.pix_off = np.concatenate([np.asarray(_pix_off, np.int64), [_total_px - 16]])
            self.word_off = np.concatenate([np.asarray(_word_off, np.int64), [_total_words]])
        ntiles = (nwords + TILE_WORDS - 1) // TILE_WORDS
        self.tile0 = np.concatenate([[0], np.cumsum(ntiles)]).astype(np.int64)
        self.ntiles = ntiles
        self.total_px = int(self.pix_off[-1]) + 16  # tail slack for vector loads
        self.pixels = int(self.npx.sum())
        self.total_words = int(self.word_off[-1])
        self.n_tiles = int(self.tile0[-1])
        if self.n_tiles >= 2 ** 31 or self.total_words >= 2 ** 40:
            raise ValueError("batch too large")
        self.max_h = int(hs.max()) if hs.size else 0
        self.max_w = int(ws.max()) if ws.size else 0
        vig = np.zeros(self.n_img, VIG_DTYPE)
        vig["pix_off"] = self.pix_off[:-1]
        vig["word_off"] = self.word_off[:-1]
        vig["h"] = hs
        vig["w"] = ws
        vig["wpr"] = wpr
        vig["tile0"] = self.tile0[:-1]
        self.vig = vig
        self._tiles = None

    @property
    def tiles(self):
        """Tile list of the per-operator kernels (built on first use: the band pipeline does not need it)."""
        if self._tiles is None:
            tiles = np.zeros(self.n_tiles, TILE_DTYPE)
            img_of_tile = np.repeat(np.arange(self.n_img, dtype=np.int64), self.ntiles)
            tiles["img"] = img_of_tile
            tiles["word0"] = (np.arange(self.n_tiles, dtype=np.int64) - self.tile0[img_of_tile]) * TILE_WORDS
            self._tiles = tiles
        return self._tiles

    def subset(self, indices):
        """Geometry of some vignettes of this batch that keeps THEIR offsets, so kernels launched on the
        subset read and write the parent's buffers in place."""
        idx = np.asarray(indices, dtype=np.int64)
        return BatchGeometry(self.h[idx], self.w[idx], _pix_off=self.pix_off[idx], _word_off=self.word_off[idx],
                             _total_px=self.total_px, _total_words=self.total_words)

    def fused_classes(self):
        """(img_list, class_off[4], leftovers): vignette indices grouped by the size classes of
        maze_vignette_stage (largest first inside a class), and the ones too large for it."""
        small = (self.h < 65536) & (self.w < 65536)  # the fused kernel keeps run coordinates in 16 bits
        # smallest class whose shared memory holds the planes AND a run list of at least two runs per row
        # (tall narrow vignettes have few words but one run per row)
        cls = np.full(self.n_img, len(FUSED_CAPS), np.int64)
        for c in range(len(FUSED_CAPS) - 1, -1, -1):
            cap = FUSED_CAPS[c]
            rc = (4 * (2 * cap - self.nwords) - 2 * (self.h + 2)) // 10
            ok = (self.nwords <= cap) & (rc >= 2 * self.h + 64) & small
            cls[ok] = c
        # the last class takes whatever fits its planes (an overflowing run list falls back at run time)
        cls[(cls == len(FUSED_CAPS)) & (self.nwords <= FUSED_CAPS[-1]) & small] = len(FUSED_CAPS) - 1
        order, off = [], [0]
        for c in range(len(FUSED_CAPS)):
            sel = np.nonzero(cls == c)[0]
            sel = sel[np.argsort(-self.nwords[sel], kind="stable")]
            order.append(sel)
            off.append(off[-1] + len(sel))
        left = np.nonzero(cls == len(FUSED_CAPS))[0]
        return np.concatenate(order).astype(np.int32), np.asarray(off, np.int32), left

    def band_plan(self, halo: int):
        """(bands, band_off, leftovers) for maze_band_stage: every vignette is cut into row bands that, with `halo`
        recomputed rows on each side, hold at most BAND_PLANE_WORDS words (a vignette that fits is one band without
        halo).  Leftovers: a side >= 65536, or so wide that a band would be mostly halo."""
        wpr = (self.w + 31) // 32
        small = (self.h < 65536) & (self.w < 65536)
        single = self.nwords <= BAND_PLANE_WORDS
        rows_max = BAND_PLANE_WORDS // wpr - 2 * halo
        ok = small & (single | (rows_max >= max(1, halo)))
        rows_max = np.maximum(rows_max, 1)
        nb = np.where(single, 1, (self.h + rows_max - 1) // rows_max)
        rpb = (self.h + nb - 1) // nb
        nb = np.where(ok, (self.h + rpb - 1) // rpb, 0)
        # a batch that is one frame (or a few): full-height bands would leave most SMs without a CTA (a 4096 x 4096
        # frame is ~100 bands for 148 SMs x 4 slots).  Frames -- which are labelled in global memory, whatever their
        # number of bands -- are cut into thinner bands, down to twice the halo, until the batch has ~3 CTAs per SM.
        frames = ok & (self.npx >= HUGE_PX) & (nb > 1)
        total = int(nb.sum())
        if frames.any() and total < BAND_TARGET_CTAS:
            thin = np.maximum(np.ceil(rpb * (total / BAND_TARGET_CTAS)).astype(np.int64), max(2 * halo, 4))
            rpb = np.where(frames, np.minimum(rpb, thin), rpb)
            nb = np.where(ok, (self.h + rpb - 1) // rpb, 0)
        band_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        n_bands = int(band_off[-1])
        if n_bands >= 2 ** 31:
            raise ValueError("batch too large")
        bands = np.zeros(n_bands, BAND_DTYPE)
        img = np.repeat(np.arange(self.n_img, dtype=np.int64), nb)
        j = np.arange(n_bands, dtype=np.int64) - band_off[img]
        bands["img"] = img
        bands["y0"] = j * rpb[img]
        bands["y1"] = np.minimum(self.h[img], (j + 1) * rpb[img])
        bands["rpb"] = rpb[img]
        return bands, band_off.astype(np.int32), np.nonzero(~ok)[0]

    @classmethod
    def from_images(cls, images):
        return cls([im.shape[0] for im in images], [im.shape[1] for im in images])

    def view(self, flat, i):
        """(h, w) view of vignette i inside a flat per-pixel host array."""
        o = int(self.pix_off[i])
        return flat[o:o + int(self.npx[i])].reshape(int(self.h[i]), int(self.w[i]))

    def pack_host(self, images, out=None, dtype=np.uint8, threads=None):
        """Copy a list of (h, w) arrays into one flat host array laid out like the device batch
        (multi-threaded memcpy in libmaze_b200.so when the arrays already have the right dtype)."""
        out, arrs, ptrs, nbytes, offs, threads = self._pack_args(images, out, dtype, threads)
        if len(arrs) == 0:
            return out
        check(lib().maze_host_pack(ptrs.ctypes.data, nbytes.ctypes.data, offs.ctypes.data, len(arrs),
                                   out.__array_interface__["data"][0], int(threads)), "maze_host_pack")
        return out

    def _pack_args(self, images, out, dtype, threads):
        dtype = np.dtype(dtype)
        if out is None:
            out = np.zeros(self.total_px, dtype)
        arrs = []
        for im in images:
            a = np.asarray(im)
            if a.dtype != dtype or not a.flags.c_contiguous:
                a = np.ascontiguousarray(a, dtype=dtype)
            arrs.append(a)
        n = len(arrs)
        ptrs = np.fromiter((a.__array_interface__["data"][0] for a in arrs), dtype=np.uint64, count=n)
        nbytes = (self.npx * dtype.itemsize).astype(np.int64)
        offs = (self.pix_off[:-1] * dtype.itemsize).astype(np.int64)
        if threads is None:
            env = os.environ.get("MAZE_PACK_THREADS")
            try:
                avail = len(os.sched_getaffinity(0))
            except Exception:
                avail = os.cpu_count() or 2
            # ranks of one box share its cores (torchrun sets LOCAL_WORLD_SIZE): eight ranks with sixteen copy threads
            # each oversubscribe a 32-core host (measured at N = 8: 478 k vignettes/s end to end with 16 threads per rank,
            # 502 k with 4)
            lws = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
            threads = int(env) if env else (min(16, max(1, avail // 2)) if lws == 1 else min(8, max(2, avail // lws)))
        return out, arrs, ptrs, nbytes, offs, threads

    def pack_host_start(self, images, out, dtype=np.uint8, threads=None):
        """pack_host that returns at once: the copy threads run in the background; call the returned function to
        wait for them (it keeps the source arrays alive until then)."""
        out, arrs, ptrs, nbytes, offs, threads = self._pack_args(images, out, dtype, threads)
        job = lib().maze_host_pack_start(ptrs.ctypes.data, nbytes.ctypes.data, offs.ctypes.data, len(arrs),
                                         out.__array_interface__["data"][0], int(threads))
        if not job:
            raise ValueError("maze_host_pack_start: bad argument")
        keep = [arrs, out]

        def wait():
            if keep:
                check(lib().maze_host_pack_wait(job), "maze_host_pack_wait")
                keep.clear()
        return wait


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def fold_erosion_radius(radius) -> int:
    """max{k : sqrt_f64(k) <= radius} (or -1): ``dist > radius`` <=> ``d2 > k`` (isotropic.py:36)."""
    r = float(radius)
    if math.isnan(r):
        return _INT_MAX  # dist > nan is False everywhere
    if r < 0:
        return -1
    if r >= 46340.0:
        return _INT_MAX
    k = int(math.floor(r * r))
    while math.sqrt(k + 1) <= r:
        k += 1
    while k >= 0 and math.sqrt(k) > r:
        k -= 1
    return k


def fold_dilation_radius(radius) -> int:
    """max{k : sqrt_f64(k) < radius} (or -1): ``dist < radius`` <=> ``d2 <= k`` (isotropic.py:67)."""
    r = float(radius)
    if math.isnan(r) or r <= 0:
        return -1
    if r >= 46340.0:
        return _INT_MAX
    k = int(math.floor(r * r))
    while math.sqrt(k + 1) < r:
        k += 1
    while k >= 0 and not (math.sqrt(k) < r):
        k -= 1
    return k


def fold_threshold(thr) -> int:
    """uint8 pixel > thr  <=>  pixel > floor(thr) (loki/pipeline.py:649, float threshold)."""
    t = float(thr)
    if math.isnan(t):
        return 255  # px > nan is False
    return int(min(255, max(-1, math.floor(t))))


_ITEMSIZE = {torch.uint8: 1, torch.int32: 4, torch.int64: 8, torch.float64: 8, torch.float32: 4, torch.int16: 2}


class Arena:
    """Bump allocator over one device buffer, reset once per step: the per-operator kernels need a dozen
    scratch arrays per call and a cudaMalloc inside a step costs milliseconds.  Grows (once) on demand."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = None
        self.off = 0
        self.need = 0

    def reserve(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.off = 0

    def reset(self):
        if self.buf is None or self.need > self.buf.numel():
            self.reserve(int(self.need * 1.2) + (1 << 20))
        self.off = 0
        self.need = 0

    def take(self, n, dtype):
        nbytes = (int(n) * _ITEMSIZE[dtype] + 255) // 256 * 256
        self.need += nbytes
        if self.buf is None or self.off + nbytes > self.buf.numel():
            return torch.empty(int(n), dtype=dtype, device=self.device)  # overflow: plain allocation this once
        t = self.buf[self.off:self.off + nbytes].view(dtype)[:int(n)]
        self.off += nbytes
        return t


class DeviceBatch:
    """A packed batch resident on one GPU plus the operators that act on it."""
    arena = None  # optional Arena for the scratch / output arrays of the per-operator kernels

    def __init__(self, geometry: BatchGeometry, device=None):
        if not torch.cuda.is_available():
            raise _lib.MazeLibraryError("maze_b200 needs a CUDA device (there is no CPU fallback)")
        lib()
        self.g = geometry
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._d_vig = self._d_tiles = None

    @property
    def d_vig(self):
        if self._d_vig is None:
            with torch.cuda.device(self.device):
                self._d_vig = torch.from_numpy(self.g.vig.view(np.uint8).copy()).to(self.device, non_blocking=False)
        return self._d_vig

    @property
    def d_tiles(self):
        """Tile list on the device (uploaded on first use: only the per-operator kernels need it)."""
        if self._d_tiles is None:
            with torch.cuda.device(self.device):
                self._d_tiles = torch.from_numpy(self.g.tiles.view(np.uint8).copy()).to(self.device, non_blocking=False)
        return self._d_tiles

    def upload_descriptors(self, pool, halo=None):
        """Vignette descriptors and (halo given) the band plan through ONE pinned staging buffer and ONE asynchronous
        copy on the current stream, instead of a blocking upload per array (each of which waits for everything queued
        on the stream, e.g. the image upload of the previous batch).  pool: a stage._PinnedPool that is not reused
        before this batch is complete."""
        if self._d_vig is not None:
            return
        parts = [self.g.vig.view(np.uint8)]
        plan = None
        if halo is not None:
            plan = self.g.band_plan(halo)
            parts += [plan[0].view(np.uint8), plan[1].view(np.uint8)]
        offs, total = [], 0
        for p in parts:
            offs.append(total)
            total += (p.size + 255) // 256 * 256
        h = pool.get("desc", max(total, 256), torch.uint8)
        hn = h.numpy()
        for p, o in zip(parts, offs):
            hn[o:o + p.size] = p.reshape(-1)
        d = pool.dev("desc", max(total, 256), torch.uint8, self.device)
        d.copy_(h, non_blocking=True)
        self._d_vig = d[offs[0]:offs[0] + parts[0].size]
        if plan is not None:
            bands, band_off, left = plan
            d_bands = d[offs[1]:offs[1] + max(parts[1].size, 16)]
            d_off = d[offs[2]:offs[2] + parts[2].size].view(torch.int32)
            self.__dict__.setdefault("_bands", {})[halo] = (d_bands, d_off, len(bands), left, bands, band_off)

    # ---- buffers -------------------------------------------------------------------------------
    def _new(self, n, dtype):
        if self.arena is not None:
            return self.arena.take(n, dtype)
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def empty_px(self, dtype):
        return self._new(self.g.total_px, dtype)

    def empty_plane(self):
        return self._new(max(self.g.total_words, 1), torch.int32)

    def empty_flags(self):
        return self._new(max(self.g.n_img, 1), torch.int32)

    def upload(self, host_flat):
        t = torch.from_numpy(host_flat) if isinstance(host_flat, np.ndarray) else host_flat
        return t.to(self.device, non_blocking=True)

    def _geo(self):
        g = self.g
        return self.d_vig.data_ptr(), g.n_img, self.d_tiles.data_ptr(), g.n_tiles

    # ---- operators -----------------------------------------------------------------------------
    def threshold_pack(self, d_image, t_int: int):
        """loki/pipeline.py:649 -> (bit plane, flags)."""
        vig, n, tiles, nt = self._geo()
        bits, flags = self.empty_plane(), self.empty_flags()
        check(lib().maze_threshold_pack(d_image.data_ptr(), vig, n, tiles, nt, int(t_int), bits.data_ptr(),
                                        flags.data_ptr(), _stream()), "maze_threshold_pack")
        return bits, flags

    def _wide_plan(self):
        """CTA prefix sums of maze_morph_pass_wide (device int64 arrays, cached)."""
        if not hasattr(self, "_wide"):
            g = self.g
            wpr = (g.w + 31) // 32
            a = np.concatenate([[0], np.cumsum(((g.h + 31) // 32) * ((wpr + 7) // 8))]).astype(np.int64)
            b = np.concatenate([[0], np.cumsum(((g.h + 7) // 8) * ((g.w + 255) // 256))]).astype(np.int64)
            self._wide = (torch.from_numpy(a).to(self.device), int(a[-1]), torch.from_numpy(b).to(self.device), int(b[-1]))
        return self._wide

    def morph_pass_wide(self, bits, flags, t: int, invert: int):
        """maze_morph_pass for large radii (separable vertical-distance + row-test kernels)."""
        d_a, n_a, d_b, n_b = self._wide_plan()
        out, fout = self.empty_plane(), self.empty_flags()
        gbuf, cbuf = self.empty_px(torch.uint8), self.empty_plane()
        check(lib().maze_morph_pass_wide(bits.data_ptr(), out.data_ptr(), self.d_vig.data_ptr(), self.g.n_img,
                                         d_a.data_ptr(), n_a, d_b.data_ptr(), n_b, int(t), int(invert), flags.data_ptr(),
                                         fout.data_ptr(), gbuf.data_ptr(), cbuf.data_ptr(), _stream()), "maze_morph_pass_wide")
        return out, fout

    def morph_pass(self, bits, flags, t: int, invert: int):
        if t >= 0 and math.isqrt(int(t)) >= WIDE_MIN_R:
            return self.morph_pass_wide(bits, flags, t, invert)
        if t < -1:  # registered footprint
            from .morphology import pass_radius
            if pass_radius(t) >= WIDE_MIN_R:
                return self.morph_pass_wide(bits, flags, t, invert)
        vig, n, tiles, nt = self._geo()
        out, fout = self.empty_plane(), self.empty_flags()
        check(lib().maze_morph_pass(bits.data_ptr(), out.data_ptr(), vig, n, tiles, nt, int(t), int(invert),
                                    flags.data_ptr(), fout.data_ptr(), _stream()), "maze_morph_pass")
        return out, fout

    def edt_sq(self, bits, invert: int = 0):
        """Exact squared EDT of the plane (or of its complement) as per-pixel int32."""
        g = self.g
        d2 = self.empty_px(torch.int32)
        scratch = self.empty_flags()
        check(lib().maze_edt_sq(bits.data_ptr(), self.d_vig.data_ptr(), g.n_img, g.max_h, g.max_w, int(invert),
                                d2.data_ptr(), scratch.data_ptr(), _stream()), "maze_edt_sq")
        return d2

    def compare_pack(self, d2, t: int, greater: int):
        vig, n, tiles, nt = self._geo()
        bits, flags = self.empty_plane(), self.empty_flags()
        check(lib().maze_compare_pack(d2.data_ptr(), vig, n, tiles, nt, int(t), int(greater), bits.data_ptr(),
                                      flags.data_ptr(), _stream()), "maze_compare_pack")
        return bits, flags

    def erosion(self, bits, flags, radius):
        """isotropic.py:35-36."""
        t = fold_erosion_radius(radius)
        if t < _DISK_T_LIMIT:
            return self.morph_pass(bits, flags, t, 0)
        if t <= WIDE_MAX_R ** 2:
            return self.morph_pass_wide(bits, flags, t, 0)
        return self.compare_pack(self.edt_sq(bits, 0), t, 1)

    def dilation(self, bits, flags, radius):
        """isotropic.py:66-67 (strict <)."""
        t = fold_dilation_radius(radius)
        if t < _DISK_T_LIMIT:
            return self.morph_pass(bits, flags, t, 1)
        if t <= WIDE_MAX_R ** 2:
            return self.morph_pass_wide(bits, flags, t, 1)
        return self.compare_pack(self.edt_sq(bits, 1), t, 0)

    def opening(self, bits, flags, radius):
        """isotropic.py:97-98."""
        return self.dilation(*self.erosion(bits, flags, radius), radius)

    def closing(self, bits, flags, radius):
        """isotropic.py:128-129."""
        return self.erosion(*self.dilation(bits, flags, radius), radius)

    def unpack_mask(self, bits, out=None):
        vig, n, tiles, nt = self._geo()
        mask = self.empty_px(torch.uint8) if out is None else out
        check(lib().maze_unpack_mask(bits.data_ptr(), vig, n, tiles, nt, mask.data_ptr(), _stream()),
              "maze_unpack_mask")
        return mask

    def label(self, bits, labels=None):
        """loki/pipeline.py:430-433 -> (labels int32 per pixel, lab_off int32[n_img+1])."""
        vig, n, tiles, nt = self._geo()
        if labels is None:
            labels = self.empty_px(torch.int32)
        parent = self.empty_px(torch.int32)
        tile_scan = self._new(self.g.n_tiles + 1, torch.int32)
        lab_off = self._new(self.g.n_img + 1, torch.int32)
        check(lib().maze_label(bits.data_ptr(), vig, n, tiles, nt, parent.data_ptr(), labels.data_ptr(),
                               tile_scan.data_ptr(), lab_off.data_ptr(), _stream()), "maze_label")
        return labels, lab_off

    def max_label(self, labels):
        vig, n, tiles, nt = self._geo()
        out = self.empty_flags()
        check(lib().maze_max_label(labels.data_ptr(), vig, n, tiles, nt, out.data_ptr(), _stream()),
              "maze_max_label")
        return out

    def lab_off_from_bounds(self, bounds):
        """Exclusive prefix sum of per-vignette label bounds (host array) as a device lab_off."""
        off = np.concatenate([[0], np.cumsum(np.asarray(bounds, dtype=np.int64))])
        if off[-1] >= 2 ** 31:
            raise ValueError("too many labels")
        return torch.from_numpy(off.astype(np.int32)).to(self.device), int(off[-1])

    def clear_border(self, labels, lab_off, n_obj: int):
        vig, n, tiles, nt = self._geo()
        if n_obj <= 0:
            return labels
        scratch = torch.empty(n_obj, dtype=torch.int32, device=self.device)
        check(lib().maze_clear_border(labels.data_ptr(), vig, n, tiles, nt, lab_off.data_ptr(), scratch.data_ptr(),
                                      int(n_obj), _stream()), "maze_clear_border")
        return labels

    def remove_small_objects(self, labels, lab_off, n_obj: int, min_size: int):
        vig, n, tiles, nt = self._geo()
        if n_obj <= 0:
            return labels
        scratch = torch.empty(n_obj, dtype=torch.int32, device=self.device)
        check(lib().maze_remove_small_objects(labels.data_ptr(), vig, n, tiles, nt, lab_off.data_ptr(),
                                              scratch.data_ptr(), int(n_obj), int(min_size), _stream()),
              "maze_remove_small_objects")
        return labels

    def regionprops(self, lab_off, n_obj: int, labels=None, bits=None, image=None, high_order=True, runs=False,
                    table=None, acc_base=None, tiles=None):
        """Feature table (n_obj, NFEAT) float64 on the device.  runs=True: `labels` are constant along the
        word runs of `bits` (straight from label() or the label filters) -> run-based reduction."""
        vig, n, d_tiles, nt = self._geo()
        if tiles is not None:  # a device tile list restricted to some vignettes of the batch
            d_tiles, nt = tiles.data_ptr(), tiles.numel() // 8
        if table is None:
            table = torch.empty((max(n_obj, 0), NFEAT), dtype=torch.float64, device=self.device)
        if n_obj <= 0 or nt == 0:
            return table
        acc = self._new(n_obj * NACC, torch.int64)
        ext = self._new(n_obj * NEXT, torch.int32)
        check(lib().maze_regionprops(_ptr(labels), _ptr(bits), _ptr(image), vig, n, d_tiles, nt, lab_off.data_ptr(),
                                     int(n_obj), acc.data_ptr(), ext.data_ptr(), table.data_ptr(),
                                     (RP_HIGH_ORDER if high_order else 0) | (RP_RUNS if runs else 0),
                                     _ptr(acc_base), _stream()), "maze_regionprops")
        return table

    def merge_labels(self, labels, labels_out, lab_off, n_obj: int, max_distance, path_tolerance=5.0,
                     index=None, index_off=None, live=None, only=None, scratch=None):
        """merge_labels.py:29-113 for every vignette.  Returns (merge_dist, n_merge, index_state, status,
        obj_scratch) device tensors.  live: optional host array with (an upper bound of) the number of labels of
        every vignette -- vignettes with fewer than two are not touched (the reference returns at :59-60); only: the
        vignettes to work on (the others are not touched).  When only
        a few vignettes are large (MAZE_MERGE_CLUSTER_PX pixels and up) each of them gets a thread-block cluster of
        eight CTAs."""
        g = self.g
        n_obj = max(int(n_obj), 1)
        # scratch: three per-pixel int32 maps (a caller that merges batch after batch passes its own, see stage.py)
        d2a, d2b, d2c = scratch if scratch is not None else (self.empty_px(torch.int32) for _ in range(3))
        obj_scratch = torch.zeros(7 * n_obj, dtype=torch.int32, device=self.device)
        merge_dist = torch.zeros(n_obj, dtype=torch.float64, device=self.device)
        n_merge = torch.zeros(g.n_img, dtype=torch.int32, device=self.device)
        index_state = torch.zeros(2 * g.n_img, dtype=torch.int32, device=self.device)
        status = torch.zeros(g.n_img, dtype=torch.int32, device=self.device)
        have_max = max_distance is not None
        todo = np.arange(g.n_img) if only is None else np.asarray(only, np.int64)
        if live is not None and index is None and only is None:
            todo = todo[np.asarray(live)[:g.n_img] >= 2]
        # a cluster of eight CTAs finishes ONE large vignette several times sooner, but a full batch keeps every SM busy
        # with one CTA per vignette anyway (measured: 33.6 ms per 4096-vignette batch either way): clusters only when
        # the large vignettes are few (frames, small batches)
        cluster_px = int(os.environ.get("MAZE_MERGE_CLUSTER_PX", str(1 << 16)))
        big = g.npx[todo] >= cluster_px
        windowed = index is None and have_max and os.environ.get("MAZE_MERGE_WINDOWED", "1") != "0"
        if windowed or big.sum() > int(os.environ.get("MAZE_MERGE_CLUSTER_MAX", "64")):
            big[:] = False  # (the windowed kernel, maze_merge_win.cu, takes one CTA per vignette)
        for sel, cs in ((todo[~big], 1), (todo[big], 8)):
            if len(sel) == 0:
                continue
            sel = sel[np.argsort(-g.npx[sel], kind="stable")]  # largest vignettes first: they decide the makespan
            order = torch.from_numpy(sel.astype(np.int32)).to(self.device)
            check(lib().maze_merge_labels_ex(labels.data_ptr(), labels_out.data_ptr(), self.d_vig.data_ptr(), g.n_img,
                                             lab_off.data_ptr(), n_obj, _ptr(index), _ptr(index_off), int(have_max),
                                             float(max_distance) if have_max else 0.0, float(path_tolerance),
                                             d2a.data_ptr(), d2b.data_ptr(), d2c.data_ptr(), obj_scratch.data_ptr(),
                                             merge_dist.data_ptr(), n_merge.data_ptr(), index_state.data_ptr(),
                                             status.data_ptr(), order.data_ptr(), len(sel), cs, _stream()),
                  "maze_merge_labels_ex")
        return merge_dist, n_merge, index_state, status, obj_scratch

    def front_chain(self, d_src, t_int, passes, labels, mask):
        """threshold -> passes -> label -> mask with the per-operator kernels in one C call.
        Returns (final bit plane, lab_off)."""
        import ctypes
        vig, n, tiles, nt = self._geo()
        pa, pb = self.empty_plane(), self.empty_plane()
        fa, fb = self.empty_flags(), self.empty_flags()
        parent = self.empty_px(torch.int32)
        tile_scan = self._new(self.g.n_tiles + 1, torch.int32)
        lab_off = self._new(self.g.n_img + 1, torch.int32)
        pt = np.asarray([p[0] for p in passes] + [0], np.int32)
        pi = np.asarray([p[1] for p in passes] + [0], np.int32)
        final = ctypes.c_void_p(0)
        check(lib().maze_front_chain(d_src.data_ptr(), vig, n, tiles, nt, int(t_int), len(passes), pt.ctypes.data,
                                     pi.ctypes.data, pa.data_ptr(), pb.data_ptr(), fa.data_ptr(), fb.data_ptr(),
                                     parent.data_ptr(), labels.data_ptr(), tile_scan.data_ptr(), lab_off.data_ptr(),
                                     mask.data_ptr(), ctypes.addressof(final), _stream()), "maze_front_chain")
        return (pa if final.value == pa.data_ptr() else pb), lab_off

    def fused_lists(self):
        if not hasattr(self, "_fused"):
            img_list, class_off, left = self.g.fused_classes()
            self._fused = (torch.from_numpy(img_list).to(self.device), class_off, left)
        return self._fused

    def band_lists(self, halo: int):
        """Device copies of the band plan (cached per halo): (d_bands, d_band_off, n_bands, leftovers, bands_host, band_off_host)."""
        cache = self.__dict__.setdefault("_bands", {})
        if halo not in cache:
            bands, band_off, left = self.g.band_plan(halo)
            d_bands = torch.from_numpy(bands.view(np.uint8).copy()).to(self.device) if len(bands) else \
                torch.zeros(16, dtype=torch.uint8, device=self.device)
            cache[halo] = (d_bands, torch.from_numpy(band_off).to(self.device), len(bands), left, bands, band_off)
        return cache[halo]

    def tiles_of(self, indices):
        """Device tile list (full-batch vignette indices) restricted to the given vignettes (cached)."""
        key = tuple(int(i) for i in indices)
        cache = self.__dict__.setdefault("_tiles_cache", {})
        if key not in cache:
            sel = np.isin(self.g.tiles["img"], np.asarray(indices, np.int32))
            t = np.ascontiguousarray(self.g.tiles[sel])
            cache[key] = torch.from_numpy(t.view(np.uint8).copy()).to(self.device)
        return cache[key]

    def vignette_stage(self, d_image, d_intensity, t_int: int, passes, bits, mask, labels, counts, staging,
                       stage_cap, high_order=True, props=True):
        """maze_vignette_stage on every vignette that fits its size classes.  counts: int32[3*n_img] =
        n_labels | fallback | acc_base; staging = (acc, hi, ext, counter) tensors with stage_cap rows.
        Returns the indices of the vignettes that are too large for it."""
        g = self.g
        d_list, class_off, left = self.fused_lists()
        n = g.n_img
        acc, hi, ext, counter = staging
        if class_off[-1] > 0:
            pt = np.asarray([p[0] for p in passes] + [0] * (4 - len(passes)), np.int32)
            pi = np.asarray([p[1] for p in passes] + [0] * (4 - len(passes)), np.int32)
            flags = (RP_HIGH_ORDER if high_order else 0) | (0 if props else FUSED_NO_PROPS)
            check(lib().maze_vignette_stage(
                d_image.data_ptr(), _ptr(d_intensity), self.d_vig.data_ptr(), d_list.data_ptr(), class_off.ctypes.data,
                int(t_int), len(passes), pt.ctypes.data, pi.ctypes.data, flags, bits.data_ptr(), mask.data_ptr(),
                labels.data_ptr(), counts.data_ptr(), counts[n:].data_ptr(), counts[2 * n:].data_ptr(),
                counter.data_ptr(), int(stage_cap), acc.data_ptr(), hi.data_ptr(), ext.data_ptr(), _stream()),
                "maze_vignette_stage")
        return left

    def props_finish_staged(self, staging, acc_base, lab_off, n_obj, has_intensity, high_order, table):
        acc, hi, ext = staging[:3]
        check(lib().maze_props_finish_staged(acc.data_ptr(), hi.data_ptr(), ext.data_ptr(), acc_base.data_ptr(),
                                             lab_off.data_ptr(), self.g.n_img, int(n_obj), int(has_intensity),
                                             RP_HIGH_ORDER if high_order else 0, table.data_ptr(), _stream()),
              "maze_props_finish_staged")
        return table

    _shape_pool = None  # scratch of label_shape, shared by the batches of one device (grown on demand)

    def label_shape(self, table, labels=None, bits=None, out=None, pool_bytes=1 << 30, runs=False):
        """perimeter / filled_area / euler_number per row of the finished feature table (maze_label_shape): the
        RegionProperties values CalculateZooProcessFeatures reads besides the moments (loki/pipeline.py:625, 654).
        runs=True: labels are constant along the runs of bits (output of label() / the fused kernel, also after the
        label filters, NOT after merge_labels): the object planes are cut from the bit plane.
        Returns an (n_obj, NSHAPE) float64 tensor."""
        from ._lib import NSHAPE
        n_obj = int(table.shape[0])
        shape = torch.empty((n_obj, NSHAPE), dtype=torch.float64, device=self.device) if out is None else out
        if n_obj == 0:
            return shape
        g = self.g
        slab_words = 2 * (int(g.h.max()) + 2) * ((int(g.w.max()) + 2 + 31) // 32)
        n_slabs = max(1, min(n_obj, 148, pool_bytes // (4 * slab_words)))  # one 512-thread CTA per SM
        need = n_slabs * slab_words + 64
        key = (str(self.device), _stream())  # one pool per stream: calls on one stream are ordered
        pools = DeviceBatch._shape_pool = DeviceBatch._shape_pool or {}
        if key not in pools or pools[key].numel() < need:
            pools[key] = torch.empty(need, dtype=torch.int32, device=self.device)
        pool = pools[key]
        counter = pool[n_slabs * slab_words:]  # one int32 behind the slabs
        check(lib().maze_label_shape(None if labels is None else labels.data_ptr(),
                                     None if bits is None else bits.data_ptr(), self.d_vig.data_ptr(),
                                     table.data_ptr(), n_obj, pool.data_ptr(), slab_words, n_slabs, int(g.h.max()),
                                     1 if (runs and labels is not None and bits is not None) else 0, counter.data_ptr(),
                                     shape.data_ptr(), _stream()), "maze_label_shape")
        return shape

    def count_scan(self, n_labels, out=None):
        lab_off = torch.empty(self.g.n_img + 1, dtype=torch.int32, device=self.device) if out is None else out
        check(lib().maze_count_scan(n_labels.data_ptr(), self.g.n_img, lab_off.data_ptr(), _stream()),
              "maze_count_scan")
        return lab_off

    def synth(self, seed: int, img_index0: int = 0, out=None):
        vig, n, tiles, nt = self._geo()
        img = self.empty_px(torch.uint8) if out is None else out
        check(lib().maze_synth_vignettes(img.data_ptr(), vig, n, tiles, nt, int(seed), int(img_index0), _stream()),
              "maze_synth_vignettes")
        return img
