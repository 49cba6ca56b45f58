"""Generate the golden fixtures in tests/golden/ from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so the pins
are manufactured here by importing the reference's own files by path
(``maze_ipp/isotropic.py``, ``maze_ipp/merge_labels.py``, ``maze_ipp/loki/config_schema.py``)
and running them with the scipy/numpy in this image.  scikit-image and morphocut are absent,
so the regionprops pins come from OpenCV (``cv2.moments`` / ``cv2.HuMoments`` /
``cv2.connectedComponentsWithStats``), not from skimage -- those rows stay "parity unpinned"
with respect to skimage itself.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import scipy
import scipy.ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _imp(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


iso = _imp("ref_isotropic", os.path.join(REF, "maze_ipp/isotropic.py"))
ml = _imp("ref_merge_labels", os.path.join(REF, "maze_ipp/merge_labels.py"))

from maze_image_processing_pipeline_b200.synth import synth_batch  # noqa: E402

RADII = [0, 0.5, 1, 1.5, 2, 2.5, 3, 5, 8, 13, 32]


def make_masks():
    rng = np.random.default_rng(1234)
    masks = {}
    for k, img in enumerate(synth_batch(11, 4, size=(72, 90))):
        masks[f"blob{k}"] = img > 40
    for k, img in enumerate(synth_batch(12, 3, lo=20, hi=70)):
        masks[f"var{k}"] = img > 40
    masks["noise30"] = rng.random((40, 53)) < 0.3
    masks["noise70"] = rng.random((37, 64)) < 0.7
    masks["zeros"] = np.zeros((9, 11), bool)
    masks["ones"] = np.ones((9, 11), bool)
    masks["ones_big"] = np.ones((40, 70), bool)
    masks["single"] = np.zeros((15, 15), bool)
    masks["single"][7, 7] = True
    masks["hole"] = np.ones((15, 17), bool)
    masks["hole"][6, 9] = False
    masks["row"] = rng.random((1, 45)) < 0.6
    masks["col"] = rng.random((45, 1)) < 0.6
    masks["row_ones"] = np.ones((1, 33), bool)
    masks["col_zeros"] = np.zeros((33, 1), bool)
    masks["px1"] = np.ones((1, 1), bool)
    masks["px0"] = np.zeros((1, 1), bool)
    return masks


def golden_isotropic():
    out = {}
    masks = make_masks()
    for name, m in masks.items():
        out[f"{name}/input"] = np.packbits(m)
        out[f"{name}/shape"] = np.array(m.shape)
        for r in RADII:
            for op in ("erosion", "dilation", "opening", "closing"):
                res = getattr(iso, f"isotropic_{op}")(m, r)
                assert res.dtype == bool
                out[f"{name}/{op}/{r}"] = np.packbits(res)
    # uint8 0/255 input behaves like bool; out= is written and returned
    m = masks["blob0"]
    buf = np.zeros(m.shape, bool)
    ret = iso.isotropic_opening(m.astype(np.uint8) * 255, 2, out=buf)
    assert ret is buf and np.array_equal(buf, iso.isotropic_opening(m, 2))
    np.savez_compressed(os.path.join(HERE, "isotropic.npz"), **out)
    return len(out)


def golden_labels():
    out = {}
    rng = np.random.default_rng(99)
    cases = {}
    for k, img in enumerate(synth_batch(21, 4, size=(64, 80))):
        cases[f"blob{k}"] = img > 40
    cases["noise45"] = rng.random((50, 67)) < 0.45
    cases["noise60"] = rng.random((33, 129)) < 0.60
    cases["diag"] = np.eye(12, dtype=bool)
    cases["anti"] = np.eye(12, dtype=bool)[::-1].copy()
    cases["checker"] = (np.indices((9, 13)).sum(0) % 2).astype(bool)
    cases["empty"] = np.zeros((7, 9), bool)
    cases["full"] = np.ones((7, 9), bool)
    cases["wide"] = rng.random((3, 200)) < 0.5
    cases["tall"] = rng.random((200, 3)) < 0.5
    spiral = np.zeros((21, 21), bool)
    for k in range(0, 10, 2):
        spiral[k, k:21 - k] = True
        spiral[k:21 - k, 20 - k] = True
        spiral[20 - k, k:21 - k] = True
        spiral[k + 2:21 - k, k] = True
    cases["spiral"] = spiral
    for name, m in cases.items():
        lab, n = ndi.label(m, structure=np.ones((3, 3)))
        assert lab.dtype == np.int32
        out[f"{name}/input"] = np.packbits(m)
        out[f"{name}/shape"] = np.array(m.shape)
        out[f"{name}/labels"] = lab.astype(np.int32)
        out[f"{name}/n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **out)
    return len(out)


def _two_blocks():
    """SURVEY.md section 0.4 layouts: three labels in a row with small gaps."""
    lab = np.zeros((15, 30), np.int32)
    lab[5:10, 2:8] = 1
    lab[5:10, 11:14] = 2   # 3 px gap to label 1
    lab[5:10, 17:24] = 3   # 3 px gap to label 2
    return lab


def _swallow():
    """label 2 is a single pixel above the 3 px gap between blocks 1 and 3.
    It is farther from the seed (4.47) than block 3 (4.0), so 3 is bridged first, and the bridge
    (sum <= 4 + 5) covers the pixel: the aliased call then finds no label 2 and raises TypeError."""
    lab = np.zeros((15, 30), np.int32)
    lab[4:11, 2:8] = 1
    lab[0, 9] = 2
    lab[4:11, 11:20] = 3
    return lab


def golden_merge():
    out = {}
    meta = {}
    cases = {"blocks": _two_blocks(), "swallow": _swallow()}
    rng = np.random.default_rng(5)
    for k, img in enumerate(synth_batch(31, 5, size=(60, 75))):
        m = iso.isotropic_closing(iso.isotropic_opening(img > 40, 1), 2)
        cases[f"blob{k}"] = ndi.label(m, structure=np.ones((3, 3)))[0].astype(np.int32)
    sparse = np.zeros((48, 64), np.int32)
    for l in range(1, 9):
        y, x = rng.integers(2, 44), rng.integers(2, 60)
        sparse[y:y + 3, x:x + 3] = l
    cases["sparse"] = sparse
    cases["one"] = (np.arange(20 * 20).reshape(20, 20) % 7 == 0).astype(np.int32) * 0
    cases["one"][3:6, 3:6] = 4
    for name, lab in cases.items():
        out[f"{name}/input"] = lab
        for md in (None, 3, 5, 6, 10, 12.5):
            for alias in (True, False):
                for tol in (5, 0, 1.5):
                    key = f"{name}/md={md}/alias={int(alias)}/tol={tol}"
                    work = lab.copy()
                    try:
                        res, dists = ml.merge_labels(work, max_distance=md, path_tolerance=tol,
                                                     return_merge_distances=True,
                                                     labels_out=work if alias else None)
                        out[key + "/labels"] = np.asarray(res, np.int32)
                        out[key + "/dists"] = np.asarray(dists, np.float64)
                        meta[key] = {"raises": None, "identity": bool(res is work)}
                    except TypeError as e:
                        meta[key] = {"raises": "TypeError", "msg": str(e)}
    # explicit index order (user-supplied list, not sorted)
    lab = cases["sparse"]
    for idx in ([5, 1, 3, 8], [8, 7, 6, 5, 4, 3, 2, 1], [2, 99]):
        key = f"sparse/index={'-'.join(map(str, idx))}"
        work = lab.copy()
        try:
            res, dists = ml.merge_labels(work, index=list(idx), max_distance=20, return_merge_distances=True, labels_out=work)
            out[key + "/labels"] = np.asarray(res, np.int32)
            out[key + "/dists"] = np.asarray(dists, np.float64)
            meta[key] = {"raises": None}
        except TypeError as e:
            meta[key] = {"raises": "TypeError", "msg": str(e)}
    n_raise = sum(1 for v in meta.values() if v["raises"])
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "merge_labels.npz"), **out)
    return len(out), n_raise


def golden_regionprops():
    """OpenCV pins for the float features (skimage is absent).  cv2 works in (x, y) so nu/hu are
    computed on the TRANSPOSED crop to land on skimage's [row, col] indexing."""
    import cv2

    out = {}
    cases = {}
    yy, xx = np.mgrid[0:80, 0:100]
    cases["ellipse"] = ((((yy - 40) * np.cos(0.5) + (xx - 50) * np.sin(0.5)) / 30.0) ** 2
                        + ((-(yy - 40) * np.sin(0.5) + (xx - 50) * np.cos(0.5)) / 12.0) ** 2 <= 1)
    cases["square"] = np.zeros((40, 40), bool)
    cases["square"][10:30, 8:28] = True
    cases["pixel"] = np.zeros((9, 9), bool)
    cases["pixel"][4, 5] = True
    cases["hline"] = np.zeros((9, 30), bool)
    cases["hline"][4, 3:25] = True
    img = synth_batch(41, 1, size=(90, 120))[0]
    cases["blobs"] = iso.isotropic_closing(iso.isotropic_opening(img > 40, 1), 2)
    rng = np.random.default_rng(8)
    for name, m in cases.items():
        lab, n = ndi.label(m, structure=np.ones((3, 3)))
        lab = lab.astype(np.int32)
        inten = rng.integers(0, 256, size=m.shape, dtype=np.uint8)
        inten[rng.random(m.shape) < 0.1] = 0
        rows = []
        for l in range(1, n + 1):
            sel = lab == l
            ys, xs = np.nonzero(sel)
            crop = sel[ys.min():ys.max() + 1, xs.min():xs.max() + 1].astype(np.uint8)
            mom = cv2.moments(np.ascontiguousarray(crop.T), binaryImage=True)  # x<->row, y<->col
            hu = cv2.HuMoments(mom).ravel()
            vals = inten[sel]
            rows.append([
                l, mom["m00"], ys.min(), xs.min(), ys.max() + 1, xs.max() + 1,
                ys.mean(), xs.mean(),
                mom["mu20"], mom["mu11"], mom["mu02"], mom["mu30"], mom["mu21"], mom["mu12"], mom["mu03"],
                mom["nu20"], mom["nu11"], mom["nu02"], mom["nu30"], mom["nu21"], mom["nu12"], mom["nu03"],
                *hu, vals.min(), vals.max(), vals.mean(), (vals == 0).mean(),
            ])
        out[f"{name}/labels"] = lab
        out[f"{name}/intensity"] = inten
        out[f"{name}/cv2"] = np.asarray(rows, np.float64).reshape(n, -1)
    out["columns"] = np.frombuffer(json.dumps([
        "label", "area", "bbox0", "bbox1", "bbox2", "bbox3", "centroid_r", "centroid_c",
        "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03",
        "nu20", "nu11", "nu02", "nu30", "nu21", "nu12", "nu03",
        "hu0", "hu1", "hu2", "hu3", "hu4", "hu5", "hu6", "imin", "imax", "imean", "frac0"]).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "regionprops_cv2.npz"), **out)
    return len(out)


def golden_schema():
    """Field names / defaults of the two config objects the stage consumes
    (loki/config_schema.py:8-37), so the stage can be checked without the reference present."""
    sys.path.insert(0, REF)
    from maze_ipp.loki.config_schema import SegmentationPostprocessingConfig, ThresholdSegmentationConfig

    def fields(model):
        return {k: {"default": (None if v.is_required() else v.default), "type": str(v.annotation)}
                for k, v in model.model_fields.items()}

    doc = {
        "SegmentationPostprocessingConfig": fields(SegmentationPostprocessingConfig),
        "ThresholdSegmentationConfig": fields(ThresholdSegmentationConfig),
        "short_forms": {
            "threshold: 30": ThresholdSegmentationConfig.model_validate(30).model_dump(),
            "postprocess: true": SegmentationPostprocessingConfig.model_validate(True).model_dump(),
        },
    }
    with open(os.path.join(HERE, "config_schema.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    return doc


if __name__ == "__main__":
    print("numpy", np.__version__, "scipy", scipy.__version__)
    print("isotropic entries:", golden_isotropic())
    print("labels entries:", golden_labels())
    print("merge entries / raising cases:", golden_merge())
    print("regionprops entries:", golden_regionprops())
    print("schema:", list(golden_schema()))
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump({"numpy": np.__version__, "scipy": scipy.__version__, "reference": REF}, f)
