"""Pins the CPU oracle (oracle/) against the golden vectors generated from the REAL reference
(tests/golden/make_golden.py imported maze_ipp/isotropic.py and maze_ipp/merge_labels.py from
/root/reference; labels come from scipy.ndimage.label, what skimage.measure.label calls for bool
input; the regionprops pins are OpenCV moments).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import scipy_chain
from conftest import ISO_OPS, ISO_RADII, iso_cases, merge_case_args, unpack, unpack_as


def test_isotropic_c_oracle_matches_reference_golden(golden_isotropic):
    z = golden_isotropic
    n = 0
    for name in iso_cases(z):
        m = unpack(z, name)
        for r in ISO_RADII:
            for op in ISO_OPS:
                want = unpack_as(z, f"{name}/{op}/{r}", m.shape)
                got = getattr(oracle, f"isotropic_{op}")(m, r)
                assert np.array_equal(got, want), (name, op, r)
                n += 1
    assert n == 880


def test_isotropic_scipy_chain_matches_reference_golden(golden_isotropic):
    z = golden_isotropic
    for name in iso_cases(z):
        m = unpack(z, name)
        for r in (0.5, 1, 2, 5):
            for op in ISO_OPS:
                want = unpack_as(z, f"{name}/{op}/{r}", m.shape)
                assert np.array_equal(getattr(scipy_chain, op)(m, r), want), (name, op, r)


def test_edt_fast_equals_bruteforce():
    rng = np.random.default_rng(3)
    for shape, p in [((17, 23), 0.7), ((9, 40), 0.95), ((30, 5), 0.5), ((6, 6), 1.0), ((1, 9), 1.0)]:
        m = rng.random(shape) < p
        assert np.array_equal(oracle.edt_sq(m), oracle.edt_sq(m, bruteforce=True))


def test_label_oracle_matches_scipy_golden(golden_labels):
    z = golden_labels
    for name in iso_cases(z):
        m = unpack(z, name)
        lab, n = oracle.label(m)
        assert n == int(z[f"{name}/n"])
        assert lab.dtype == np.int32
        assert np.array_equal(lab, z[f"{name}/labels"]), name


def _run_merge(fn, lab, kw, alias):
    work = lab.copy()
    md = None if kw["md"] == "None" else float(kw["md"])
    return work, fn(work, max_distance=md, path_tolerance=float(kw["tol"]), return_merge_distances=True,
                    labels_out=work if alias else None)


@pytest.mark.parametrize("impl", ["c", "scipy"])
def test_merge_labels_oracle_matches_reference_golden(golden_merge, impl):
    z, meta = golden_merge
    fn = oracle.merge_labels if impl == "c" else scipy_chain.merge_labels
    checked = raised = 0
    for key, info in meta.items():
        name, kw = merge_case_args(key)
        lab = z[f"{name}/input"]
        if "index" in kw:
            idx = [int(v) for v in kw["index"].split("-")]
            work = lab.copy()
            call = lambda: fn(work, index=list(idx), max_distance=20, return_merge_distances=True, labels_out=work)
        else:
            alias = kw["alias"] == "1"
            call = lambda: _run_merge(fn, lab, kw, alias)[1]
        if info["raises"]:
            with pytest.raises(TypeError):
                call()
            raised += 1
            continue
        res, dists = call()
        assert np.array_equal(np.asarray(res), z[key + "/labels"]), key
        assert np.array_equal(np.asarray(dists, np.float64), z[key + "/dists"]), key  # bit-exact float64
        checked += 1
    assert checked > 300 and raised == 5


def test_merge_identity_when_fewer_than_two_labels():
    lab = np.zeros((8, 8), np.int32)
    lab[2:4, 2:4] = 3
    assert oracle.merge_labels(lab, max_distance=5) is lab
    assert scipy_chain.merge_labels(lab, max_distance=5) is lab


def test_regionprops_oracle_vs_opencv(golden_props):
    z, cols = golden_props
    ci = {c: i for i, c in enumerate(cols)}
    for name in ("ellipse", "square", "pixel", "hline", "blobs"):
        lab, inten, cv = z[f"{name}/labels"], z[f"{name}/intensity"], z[f"{name}/cv2"]
        t = oracle.regionprops_table(lab, inten)
        assert t.shape[0] == cv.shape[0]
        for row, ref in zip(t, cv):
            area = ref[ci["area"]]
            assert row[oracle.F_AREA] == area
            assert list(row[oracle.F_BBOX:oracle.F_BBOX + 4]) == [ref[ci[f"bbox{k}"]] for k in range(4)]
            np.testing.assert_allclose(row[oracle.F_CENTROID:oracle.F_CENTROID + 2],
                                       [ref[ci["centroid_r"]], ref[ci["centroid_c"]]], rtol=1e-12)
            mu = row[oracle.F_MU:oracle.F_MU + 16].reshape(4, 4)
            nu = row[oracle.F_NU:oracle.F_NU + 16].reshape(4, 4)
            for key, (p, q) in {"20": (2, 0), "11": (1, 1), "02": (0, 2), "30": (3, 0), "21": (2, 1),
                                "12": (1, 2), "03": (0, 3)}.items():
                scale = area ** ((p + q) / 2 + 1)
                assert abs(mu[p, q] - ref[ci["mu" + key]]) <= 1e-9 * scale, (name, key)
                assert abs(nu[p, q] - ref[ci["nu" + key]]) <= 1e-9, (name, key)
            hu = row[oracle.F_HU:oracle.F_HU + 7]
            ref_hu = np.array([ref[ci[f"hu{k}"]] for k in range(7)])
            # cv2 ran on the transposed crop, so its (x, y) powers are skimage's (row, col) powers and
            # hu[6] already carries skimage's sign (it is the opposite of cv2 on the untransposed crop)
            np.testing.assert_allclose(hu, ref_hu, rtol=1e-6, atol=1e-12)
            np.testing.assert_allclose(row[oracle.F_IMIN:oracle.F_IMIN + 4],
                                       [ref[ci["imin"]], ref[ci["imax"]], ref[ci["imean"]], ref[ci["frac0"]]],
                                       rtol=1e-12)


def test_regionprops_known_answers():
    # single pixel: zero inertia -> eccentricity 0, axes 0
    lab = np.zeros((5, 5), np.int32)
    lab[2, 3] = 1
    t = oracle.regionprops_table(lab)[0]
    assert t[oracle.F_ECC] == 0 and t[oracle.F_AXIS_MAJOR] == 0 and t[oracle.F_AREA] == 1
    # square: T00 == T11 -> the +-pi/4 branch (T01 == 0 -> -pi/4)
    lab = np.zeros((12, 12), np.int32)
    lab[2:8, 3:9] = 1
    t = oracle.regionprops_table(lab)[0]
    assert t[oracle.F_ORIENT] == -np.pi / 4 and abs(t[oracle.F_ECC]) < 1e-7
    # horizontal line of length n: major axis 4*sqrt((n^2-1)/12), orientation +-pi/2
    lab = np.zeros((5, 40), np.int32)
    lab[2, 5:30] = 1
    t = oracle.regionprops_table(lab)[0]
    np.testing.assert_allclose(t[oracle.F_AXIS_MAJOR], 4 * np.sqrt((25 ** 2 - 1) / 12), rtol=1e-12)
    assert abs(abs(t[oracle.F_ORIENT]) - np.pi / 2) < 1e-12 and t[oracle.F_ECC] == 1.0


def test_label_filters_oracle():
    lab = np.zeros((10, 12), np.int32)
    lab[0:2, 0:2] = 1      # touches border
    lab[4:6, 4:7] = 2      # 6 px
    lab[8, 8] = 3          # 1 px
    a = oracle.clear_border(lab.copy())
    assert set(np.unique(a)) == {0, 2, 3}
    b = oracle.remove_small_objects(lab.copy(), 4)
    assert set(np.unique(b)) == {0, 1, 2}
    assert np.array_equal(scipy_chain.clear_border(lab.copy()), a)
    assert np.array_equal(scipy_chain.remove_small_objects(lab.copy(), 4), b)


def test_chain_c_oracle_equals_scipy_chain():
    from maze_image_processing_pipeline_b200.synth import synth_batch
    for k, img in enumerate(synth_batch(77, 6, lo=40, hi=120)):
        mask, labels, table = scipy_chain.loki_chain(img, 40, 1, 2, merge_segments_distance=10 if k % 2 else 0)
        m = oracle.isotropic_closing(oracle.isotropic_opening(oracle.threshold(img, 40), 1), 2)
        assert np.array_equal(m, mask)
        lab, n = oracle.label(m)
        if k % 2:
            lab = oracle.merge_labels(lab, max_distance=10, labels_out=lab)
        assert np.array_equal(lab, labels)


# ---- shape features ---------------------------------------------------
def _shape_cases():
    from scipy import ndimage as ndi
    rng = np.random.default_rng(5)
    cases = []
    for t in range(120):
        h, w = rng.integers(1, 48, 2)
        img = rng.random((h, w)) < rng.choice([0.2, 0.5, 0.8, 0.95])
        if t % 3 == 0:
            img = ndi.binary_opening(ndi.binary_dilation(img, iterations=2))
        cases.append(np.asarray(img, bool))
    return cases


def test_shape_oracle_euler_and_fill_match_independent_definitions():
    """euler_number (quad counting) == #8-connected components - #4-connected holes; filled_area == everything the
    8-connected flood of the complement from outside does not reach -- both counted with ndi.label."""
    from scipy import ndimage as ndi
    from oracle import shape as S
    eight = np.ones((3, 3))
    for img in _shape_cases():
        n8 = ndi.label(img, structure=eight)[1]
        holes4 = ndi.label(np.pad(img, 1) == 0)[1] - 1
        assert S.euler_number(img) == n8 - holes4
        comp = np.pad(img, 1) == 0
        lab, _ = ndi.label(comp, structure=eight)
        assert S.filled_area(img) == comp.size - int((lab == lab[0, 0]).sum())


def test_shape_oracle_known_answers():
    from oracle import shape as S
    one = np.ones((1, 1), bool)
    assert S.perimeter(one) == 0 and S.euler_number(one) == 1 and S.filled_area(one) == 1
    for k in (2, 3, 7, 20):                              # k x k square: 4 (k - 1)
        assert S.perimeter(np.ones((k, k), bool)) == 4 * (k - 1)
    assert S.perimeter(np.ones((5, 9), bool)) == 2 * (4 + 8)
    assert S.perimeter(np.ones((1, 10), bool)) == 8      # a line: end pixels weigh 0, inner ones 1
    diag = np.eye(6, dtype=bool)                         # diagonal line: inner pixels weigh sqrt(2)
    assert abs(S.perimeter(diag) - 4 * np.sqrt(2)) < 1e-12 and S.euler_number(diag) == 1
    ring = np.ones((7, 7), bool)
    ring[2:5, 2:5] = False
    assert S.euler_number(ring) == 0 and S.filled_area(ring) == 49
    leak = ring.copy()                                   # the hole touches the outside through a diagonal step:
    leak[0, 0] = leak[0, 1] = leak[1, 0] = False         # ... not yet
    assert S.filled_area(leak) == 46
    leak[1, 1] = False                                   # now (1,1)-(2,2) connects hole and outside 8-wise
    assert S.filled_area(leak) == int(leak.sum())
    assert S.euler_number(leak) == 0                     # the 4-connected background of euler_number still sees a hole
    two = np.zeros((5, 9), np.int32)                     # label_shape works on each label's own crop
    two[1:4, 1:4] = 1
    two[2, 2] = 0
    two[0:5, 6:9] = 3
    t = S.label_shape(two)
    assert t.shape == (3, 8) and np.isnan(t[1]).all()
    assert list(t[0, :3]) == [8.0, 9.0, 0.0] and list(t[2, :3]) == [2 * (2 + 4), 15.0, 1.0]
    assert t[0, 6] == 9 and t[2, 6] == 15
    # convex hull image: hull of the edge midpoints of the pixels, centres in or on it
    assert S.convex_area(one) == 1 and S.convex_area(np.ones((4, 7), bool)) == 28
    assert S.convex_area(diag) == 6                      # the band around a diagonal line holds no other centre
    plus = np.zeros((3, 3), bool)
    plus[1, :] = plus[:, 1] = True
    assert S.convex_area(plus) == 5                      # the corners' centres lie outside the octagon
    ell = np.zeros((6, 6), bool)
    ell[:, 0] = ell[5, :] = True                         # an L: hull = triangle over the two arms
    assert S.convex_area(ell) == 21
    assert S.convex_area(ring) == 49 and S.convex_area(np.zeros((3, 3), bool)) == 0


def _convex_area_integer(img):
    """The algorithm of maze_label_shape's hull step, restated in Python: row extremes -> leftmost / rightmost
    candidate per half-row level in doubled coordinates -> monotone chains -> exact per-row counts."""
    h, w = img.shape
    ext = []
    for y in range(h):
        xs = np.nonzero(img[y])[0]
        ext.append((int(xs[0]), int(xs[-1])) if len(xs) else None)
    L, R = [], []
    for Y in range(-1, 2 * (h - 1) + 2):       # row y <-> Y = 2y; odd levels lie between two rows
        if Y % 2 == 0:
            e = ext[Y // 2]
            if e is not None:
                L.append((Y, 2 * e[0] - 1)); R.append((Y, 2 * e[1] + 1))
        else:
            c = [ext[y] for y in ((Y - 1) // 2, (Y + 1) // 2) if 0 <= y < h and ext[y] is not None]
            if c:
                L.append((Y, min(2 * e[0] for e in c))); R.append((Y, max(2 * e[1] for e in c)))

    def chain(pts, left):
        st = []
        for C in pts:
            while len(st) >= 2:
                A, B = st[-2], st[-1]
                cr = (B[1] - A[1]) * (C[0] - A[0]) - (C[1] - A[1]) * (B[0] - A[0])
                if (cr >= 0) if left else (cr <= 0):
                    st.pop()
                else:
                    break
            st.append(C)
        return st

    def at(ch, Y):
        for (Y1, X1), (Y2, X2) in zip(ch[:-1], ch[1:]):
            if Y1 <= Y <= Y2:
                return X1 * (Y2 - Y1) + (X2 - X1) * (Y - Y1), Y2 - Y1
        return None

    cl, cr = chain(L, True), chain(R, False)
    total = 0
    for y in range(h):
        a, b = at(cl, 2 * y), at(cr, 2 * y)
        if a is None or b is None:
            continue
        xmin, xmax = max(-((-a[0]) // (2 * a[1])), 0), min(b[0] // (2 * b[1]), w - 1)
        total += max(0, xmax - xmin + 1)
    return total


def test_convex_area_integer_algorithm_equals_qhull_oracle():
    """The exact-integer hull rasterisation the CUDA kernel implements against the Qhull-based oracle, on crops that
    include disconnected masks, lines and single pixels."""
    from oracle import shape as S
    n = 0
    for img in _shape_cases():
        if not img.any():
            continue
        ys, xs = np.nonzero(img)
        crop = img[ys.min():ys.max() + 1, xs.min():xs.max() + 1]
        assert _convex_area_integer(crop) == S.convex_area(crop)
        n += 1
    assert n > 100
