"""Seeded synthetic LOKI-shaped vignettes.

LOKI vignettes are bright objects on a dark background (hence ``threshold_brighter``,
loki/config_schema.py:35-37).  Each vignette is uint8: background N(12, 4) clipped, plus 1-6
rotated anisotropic Gaussian blobs with amplitude U(80, 220) and sigma U(2 %, 10 %) of each
side, centres in the middle 70 %.  Used by tests, ``smoke()`` and ``bench.py``; the device-side
generator for the large benchmark configurations evaluates the same model in a CUDA kernel
(``maze_synth_vignettes`` in include/maze_b200.h).
"""
from __future__ import annotations

import numpy as np

MAX_BLOBS = 6
BLOB_PARAMS = 6  # cy, cx, inv-cov a, b, c (q = a dy^2 + 2 b dy dx + c dx^2), amplitude


def synth_sizes(seed: int, n: int, lo: int = 64, hi: int = 1024):
    """H and W independently log-uniform integers in [lo, hi] (BASELINE.json configs[1])."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(np.log(lo), np.log(hi), size=(n, 2))
    hw = np.clip(np.rint(np.exp(u)).astype(np.int64), lo, hi)
    return hw[:, 0].astype(np.int32), hw[:, 1].astype(np.int32)


def blob_params(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    """(MAX_BLOBS, BLOB_PARAMS) float32 table; unused rows have amplitude 0."""
    nb = int(rng.integers(1, MAX_BLOBS + 1))
    p = np.zeros((MAX_BLOBS, BLOB_PARAMS), np.float32)
    for k in range(nb):
        cy = rng.uniform(0.15, 0.85) * h
        cx = rng.uniform(0.15, 0.85) * w
        sy = rng.uniform(0.02, 0.10) * h
        sx = rng.uniform(0.02, 0.10) * w
        th = rng.uniform(0.0, np.pi)
        amp = rng.uniform(80.0, 220.0)
        ct, st = np.cos(th), np.sin(th)
        a = 0.5 * (ct * ct / (sy * sy) + st * st / (sx * sx))
        c = 0.5 * (st * st / (sy * sy) + ct * ct / (sx * sx))
        b = 0.5 * ct * st * (1.0 / (sy * sy) - 1.0 / (sx * sx))
        p[k] = (cy, cx, a, b, c, amp)
    return p


def synth_vignette(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    p = blob_params(rng, h, w)
    img = rng.normal(12.0, 4.0, size=(h, w)).astype(np.float32)
    yy = np.arange(h, dtype=np.float32)[:, None]
    xx = np.arange(w, dtype=np.float32)[None, :]
    for cy, cx, a, b, c, amp in p:
        if amp == 0:
            continue
        dy = yy - cy
        dx = xx - cx
        img += amp * np.exp(-(a * dy * dy + 2 * b * dy * dx + c * dx * dx))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_batch(seed: int, n: int, size=None, lo: int = 64, hi: int = 1024):
    """List of n vignettes.  ``size=(h, w)`` fixes the shape (configs[0]: 256x256), otherwise
    shapes come from :func:`synth_sizes`."""
    rng = np.random.default_rng(seed)
    if size is None:
        hs, ws = synth_sizes(seed + 7919, n, lo, hi)
    else:
        hs = np.full(n, size[0], np.int32)
        ws = np.full(n, size[1], np.int32)
    return [synth_vignette(rng, int(h), int(w)) for h, w in zip(hs, ws)]


def synth_dense_frame(seed: int, size: int = 4096, n_blobs: int = 3000) -> np.ndarray:
    """configs[3]: one dense frame with thousands of small blobs (sigma 4-12 px)."""
    rng = np.random.default_rng(seed)
    img = rng.normal(12.0, 4.0, size=(size, size)).astype(np.float32)
    for _ in range(n_blobs):
        cy, cx = rng.uniform(0, size, 2)
        s = rng.uniform(4.0, 12.0)
        amp = rng.uniform(80.0, 220.0)
        r = int(4 * s) + 1
        y0, y1 = max(0, int(cy) - r), min(size, int(cy) + r + 1)
        x0, x1 = max(0, int(cx) - r), min(size, int(cx) + r + 1)
        yy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(cy)
        xx = np.arange(x0, x1, dtype=np.float32)[None, :] - np.float32(cx)
        img[y0:y1, x0:x1] += amp * np.exp(-(yy * yy + xx * xx) / (2 * s * s))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
