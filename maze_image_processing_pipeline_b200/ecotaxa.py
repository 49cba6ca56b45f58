"""Host-side tail of the LOKI pipeline: the EcoTaxa archive writer and ``rescale_max_intensity``
(maze_ipp/loki/pipeline.py:382-383, 1225-1236).

I/O bound and deliberately on the CPU: the writer owns a worker thread, so encoding and zipping the objects of
batch i overlap with the GPU work on batch i + 1 (``stage.stream_objects`` / ``stage.map`` never wait for it).

Restated, not executed: the reference uses morphocut's ``EcotaxaWriter(archive_fn, [(filename, image), ...], meta,
store_types=...)`` (morphocut is pinned by requirements.txt:1 but not available offline).  What is written follows
the EcoTaxa import format: a zip archive with the image files and one ``ecotaxa_export.tsv`` -- a header row with the
metadata keys plus ``img_file_name`` / ``img_rank``, an optional second row with the column types (``[f]`` for
numbers, ``[t]`` for text; ``type_header`` of config_schema.py:271-275), one row per object image.
"""
from __future__ import annotations

import io
import os
import queue
import threading
import zipfile
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def rescale_max_intensity(image: np.ndarray) -> np.ndarray:
    """loki/pipeline.py:382-383: ``skimage.exposure.rescale_intensity(image, (0, image.max()))`` -- the grey values
    are stretched so that the brightest pixel reaches the top of the dtype's range (same float64 operations and
    the same truncating cast as skimage)."""
    image = np.asarray(image)
    imin, imax = 0, image.max() if image.size else 0
    if np.issubdtype(image.dtype, np.integer):
        omin, omax = 0, np.iinfo(image.dtype).max
    else:
        omin, omax = 0.0, 1.0
    out = np.clip(image, imin, imax)
    if imin != imax:
        out = (out - imin) / (imax - imin)
        return np.asarray(out * (omax - omin) + omin, dtype=image.dtype)
    return np.clip(out, omin, omax).astype(image.dtype)


def _encode(filename: str, image: np.ndarray) -> bytes:
    from PIL import Image
    img = np.asarray(image)
    if img.dtype == bool:
        img = img.astype(np.uint8) * 255
    ext = os.path.splitext(filename)[1].lower().lstrip(".")
    fmt = {"jpg": "JPEG", "jpeg": "JPEG", "png": "PNG", "tif": "TIFF", "tiff": "TIFF", "bmp": "BMP"}.get(ext)
    if fmt is None:
        raise ValueError(f"unsupported image extension in {filename!r}")
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format=fmt)
    return buf.getvalue()


def _column_type(values: Sequence) -> str:
    numeric = all(isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, (bool, np.bool_))
                  for v in values if v is not None and v != "")
    return "[f]" if numeric else "[t]"


def _cell(v) -> str:
    if v is None:
        return ""
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


class EcotaxaWriter:
    """``with EcotaxaWriter(archive_fn, store_types=config.type_header) as w: w.add([(fn, image), ...], meta)``.

    ``add`` returns immediately; a worker thread encodes the images (format by file extension) into the zip.  On
    ``close`` (or at the end of the ``with`` block) the table is written as ``ecotaxa_export.tsv``: the columns are the
    union of all metadata keys in first-seen order, preceded by ``img_file_name`` and ``img_rank``."""

    TSV_NAME = "ecotaxa_export.tsv"

    def __init__(self, archive_fn: str, store_types: bool = True, queue_size: int = 256):
        self.archive_fn = archive_fn
        self.store_types = store_types
        self._rows: List[Dict] = []
        self._columns: List[str] = ["img_file_name", "img_rank"]
        self._q: "queue.Queue" = queue.Queue(maxsize=queue_size)
        self._error: Optional[BaseException] = None
        self._zip = zipfile.ZipFile(archive_fn, "w", compression=zipfile.ZIP_DEFLATED)
        self._thread = threading.Thread(target=self._work, name="ecotaxa-writer", daemon=True)
        self._thread.start()
        self._closed = False

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                if self._error is None:
                    fn, image = item
                    self._zip.writestr(fn, _encode(fn, image))
            except BaseException as e:  # surfaced by the next add() / close()
                self._error = e

    def add(self, images: Iterable[Tuple[str, np.ndarray]], meta: Dict):
        if self._closed:
            raise ValueError("writer is closed")
        if self._error is not None:
            raise self._error
        for rank, (fn, image) in enumerate(images, start=1):
            row = {"img_file_name": fn, "img_rank": rank}
            for k, v in meta.items():
                if k not in row:
                    row[k] = v
                if k not in self._columns:
                    self._columns.append(k)
            self._rows.append(row)
            self._q.put((fn, np.array(image, copy=True)))  # the caller's buffers are recycled by the next batch

    def close(self):
        if self._closed:
            return
        self._closed = True
        self._q.put(None)
        self._thread.join()
        try:
            if self._error is None:
                lines = ["\t".join(self._columns)]
                if self.store_types:
                    lines.append("\t".join(_column_type([r.get(c) for r in self._rows]) for c in self._columns))
                for r in self._rows:
                    lines.append("\t".join(_cell(r.get(c)) for c in self._columns))
                self._zip.writestr(self.TSV_NAME, "\n".join(lines) + "\n")
        finally:
            self._zip.close()
        if self._error is not None:
            raise self._error

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.close()
        return False
