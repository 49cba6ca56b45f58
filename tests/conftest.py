import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def unpack(z, name):
    shape = tuple(int(v) for v in z[f"{name}/shape"])
    n = shape[0] * shape[1]
    return np.unpackbits(z[f"{name}/input"])[:n].reshape(shape).astype(bool)


def unpack_as(z, key, shape):
    n = shape[0] * shape[1]
    return np.unpackbits(z[key])[:n].reshape(shape).astype(bool)


@pytest.fixture(scope="session")
def golden_isotropic():
    return np.load(os.path.join(GOLDEN, "isotropic.npz"))


@pytest.fixture(scope="session")
def golden_labels():
    return np.load(os.path.join(GOLDEN, "labels.npz"))


@pytest.fixture(scope="session")
def golden_merge():
    z = np.load(os.path.join(GOLDEN, "merge_labels.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def golden_props():
    z = np.load(os.path.join(GOLDEN, "regionprops_cv2.npz"))
    cols = json.loads(bytes(z["columns"]).decode())
    return z, cols


ISO_RADII = [0, 0.5, 1, 1.5, 2, 2.5, 3, 5, 8, 13, 32]
ISO_OPS = ("erosion", "dilation", "opening", "closing")


def iso_cases(z):
    return sorted({k.split("/")[0] for k in z.files})


def merge_case_args(key):
    """'blob0/md=5/alias=1/tol=0' -> (name, dict)"""
    parts = key.split("/")
    name = parts[0]
    kw = {}
    for p in parts[1:]:
        k, v = p.split("=")
        kw[k] = v
    return name, kw
