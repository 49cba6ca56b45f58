"""Randomised parity (hypothesis) of the CUDA path against the oracle: arbitrary small masks, radii and label
images, including degenerate shapes.  Bit-exact."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

pytestmark = pytest.mark.gpu

SET = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@st.composite
def masks(draw, max_side=70):
    h = draw(st.integers(1, max_side))
    w = draw(st.integers(1, max_side))
    density = draw(st.sampled_from([0.0, 0.05, 0.3, 0.5, 0.7, 0.95, 1.0]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    m = rng.random((h, w)) < density
    if draw(st.booleans()) and h > 4 and w > 4:  # a solid block so that erosions leave something
        y, x = draw(st.integers(0, h - 3)), draw(st.integers(0, w - 3))
        m[y:y + draw(st.integers(1, h - y)), x:x + draw(st.integers(1, w - x))] = True
    return m


@settings(**SET)
@given(m=masks(), radius=st.sampled_from([0, 0.5, 1, 1.4142135623730951, 1.5, 2, 2.2, 2.9, 3, 4.5, 7, 33.5]),
       op=st.sampled_from(["erosion", "dilation", "opening", "closing"]))
def test_isotropic_random(m, radius, op):
    from maze_image_processing_pipeline_b200 import isotropic
    want = getattr(oracle, f"isotropic_{op}")(m, radius)
    got = getattr(isotropic, f"isotropic_{op}")(m, radius)
    assert np.array_equal(got, want)


@settings(**SET)
@given(m=masks(max_side=90))
def test_label_random(m):
    from maze_image_processing_pipeline_b200 import measure
    want, n = oracle.label(m)
    got, gn = measure.label(m, return_num=True)
    assert gn == n and np.array_equal(got, want)


@settings(**SET)
@given(ms=st.lists(masks(max_side=48), min_size=1, max_size=6), r_open=st.sampled_from([0, 1, 1.5, 2, 3]),
       r_close=st.sampled_from([0, 1, 2, 2.5, 3]), thr=st.sampled_from([0, 40, 127.5, 254]))
def test_stage_random_batches(ms, r_open, r_close, thr):
    """The fused stage on batches of arbitrary small vignettes (grey values from the mask + noise)."""
    from oracle import scipy_chain
    from maze_image_processing_pipeline_b200 import stage as S
    rng = np.random.default_rng(len(ms))
    imgs = [np.where(m, rng.integers(128, 256, m.shape), rng.integers(0, 128, m.shape)).astype(np.uint8) for m in ms]
    pp = S.SegmentationPostprocessingConfig(closing_radius=r_close, opening_radius=r_open)
    res = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(thr), pp)(imgs)
    for i, im in enumerate(imgs):
        mask, labels, table = scipy_chain.loki_chain(im, thr, r_open, r_close)
        assert np.array_equal(res.mask(i), mask)
        assert np.array_equal(res.labels(i), labels)
        feats = res.features(i)
        assert len(feats) == len(table)
        assert np.array_equal(feats[:, oracle.F_AREA], table[:, oracle.F_AREA])
        assert np.array_equal(feats[:, oracle.F_BBOX:oracle.F_BBOX + 4], table[:, oracle.F_BBOX:oracle.F_BBOX + 4])
        assert np.array_equal(feats[:, oracle.F_IMIN:oracle.F_IMAX + 1], table[:, oracle.F_IMIN:oracle.F_IMAX + 1])
        np.testing.assert_allclose(feats[:, oracle.F_CENTROID:oracle.F_CENTROID + 2],
                                   table[:, oracle.F_CENTROID:oracle.F_CENTROID + 2], rtol=1e-13)


@settings(**SET)
@given(m=masks(max_side=90), as_labels=st.booleans())
def test_label_shape_random(m, as_labels):
    """perimeter class counts, euler_number and filled_area of arbitrary masks: per 8-connected label, or the whole
    mask as one region (ImageProperties semantics; the region may then be disconnected)."""
    from oracle import shape as oshape
    from maze_image_processing_pipeline_b200 import measure
    if as_labels:
        lab, _ = oracle.label(m)
        got, want = measure.regionprops_shape(lab), oshape.label_shape(lab)
    else:
        got, want = measure.mask_shape(m), oshape.label_shape(m.astype(np.int32), max_label=1)
    assert got.shape == want.shape
    absent = np.isnan(want[:, 0])
    assert np.array_equal(np.isnan(got[:, 0]), absent)
    assert np.array_equal(got[~absent, 1:7], want[~absent, 1:7])
    np.testing.assert_allclose(got[~absent, 0], want[~absent, 0], rtol=1e-12)


@settings(**SET)
@given(m=masks(max_side=80), r=st.integers(1, 14), crosses=st.booleans(),
       op=st.sampled_from(["erosion", "dilation", "opening", "closing"]))
def test_binary_morphology_random(m, r, crosses, op):
    """skimage-style binary morphology with disk(r) / disk(r, decomposition="crosses") on arbitrary masks."""
    from oracle import scipy_chain
    from maze_image_processing_pipeline_b200 import morphology as M
    fp = M.disk(r, decomposition="crosses" if crosses else None)
    want = getattr(scipy_chain, f"binary_{op}")(m, fp)
    got = getattr(M, f"binary_{op}")(m, fp)
    assert np.array_equal(got, want)
