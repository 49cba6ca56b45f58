"""Parity of the CUDA path (through the C-ABI) with the reference-pinned oracle.  Bit-exact for
masks, labels, distances (float64 merge distances included) and all integer features; float
features within rel 1e-5 (north_star) with an absolute floor scaled by area^((p+q)/2+1)."""
import os

import numpy as np
import pytest

import oracle
from oracle import scipy_chain
from conftest import ISO_OPS, ISO_RADII, iso_cases, merge_case_args, unpack, unpack_as

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mz():
    import torch
    import maze_image_processing_pipeline_b200 as pkg
    from maze_image_processing_pipeline_b200 import device, isotropic, measure, merge_labels, stage, synth, _lib
    _lib.lib()

    class NS:
        pass
    ns = NS()
    ns.torch, ns.device, ns.isotropic, ns.measure, ns.stage, ns.synth = torch, device, isotropic, measure, stage, synth
    ns.merge_labels = merge_labels.merge_labels
    return ns


# ------------------------------------------------------------------------------------------------
def test_threshold_float_thresholds(mz):
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in [(7, 33), (64, 64), (5, 100), (1, 1), (40, 31)]]
    geom = mz.device.BatchGeometry.from_images(imgs)
    batch = mz.device.DeviceBatch(geom)
    d_img = batch.upload(geom.pack_host(imgs))
    for thr in (-3, -0.5, 0, 0.5, 30, 30.5, 127.999, 254, 254.5, 255, 300.0):
        bits, flags = batch.threshold_pack(d_img, mz.device.fold_threshold(thr))
        mask = batch.unpack_mask(bits).cpu().numpy()
        fl = flags.cpu().numpy()
        for i, im in enumerate(imgs):
            want = im > thr
            assert np.array_equal(geom.view(mask, i).astype(bool), want), (thr, i)
            assert bool(fl[i] & 1) == bool(want.any()) and bool(fl[i] & 2) == bool((~want).any())


def test_isotropic_golden_per_image_api(mz, golden_isotropic):
    z = golden_isotropic
    for name in iso_cases(z):
        m = unpack(z, name)
        for r in (0, 0.5, 1, 1.5, 2, 3, 32):
            for op in ISO_OPS:
                want = unpack_as(z, f"{name}/{op}/{r}", m.shape)
                got = getattr(mz.isotropic, f"isotropic_{op}")(m, r)
                assert got.dtype == bool and np.array_equal(got, want), (name, op, r)
    # uint8 0/255 input and out= conventions (isotropic.py:36: np.greater(..., out=out) returns out)
    m = unpack(z, "blob0")
    buf = np.zeros(m.shape, bool)
    ret = mz.isotropic.isotropic_opening(m.astype(np.uint8) * 255, 2, out=buf)
    assert ret is buf and np.array_equal(buf, unpack_as(z, "blob0/opening/2", m.shape))


def test_isotropic_golden_batched_all_radii(mz, golden_isotropic):
    z = golden_isotropic
    names = iso_cases(z)
    masks = [unpack(z, n) for n in names]
    geom = mz.device.BatchGeometry.from_images(masks)
    batch = mz.device.DeviceBatch(geom)
    d = batch.upload(geom.pack_host([m.view(np.uint8) for m in masks]))
    bits, flags = batch.threshold_pack(d, 0)
    n = 0
    for r in ISO_RADII:
        for op in ISO_OPS:
            out_bits, _ = getattr(batch, op)(bits, flags, r)
            got = batch.unpack_mask(out_bits).cpu().numpy()
            for i, name in enumerate(names):
                want = unpack_as(z, f"{name}/{op}/{r}", masks[i].shape)
                assert np.array_equal(geom.view(got, i).astype(bool), want), (name, op, r)
                n += 1
    assert n == 880


def test_isotropic_large_radius_edt_path(mz):
    rng = np.random.default_rng(5)
    img = mz.synth.synth_batch(3, 1, size=(150, 190))[0]
    masks = [img > 40, rng.random((90, 140)) < 0.97, np.ones((60, 45), bool), np.zeros((50, 70), bool)]
    for m in masks:
        for r in (33, 40.5, 64, 200):
            for op in ISO_OPS:
                want = getattr(oracle, f"isotropic_{op}")(m, r)
                got = getattr(mz.isotropic, f"isotropic_{op}")(m, r)
                assert np.array_equal(got, want), (m.shape, op, r)


def test_edt_sq_exact(mz):
    rng = np.random.default_rng(6)
    cases = [rng.random((37, 53)) < 0.8, rng.random((128, 200)) < 0.995, np.ones((20, 31), bool),
             np.ones((1, 17), bool), np.ones((17, 1), bool), np.zeros((4, 4), bool), rng.random((1, 64)) < 0.7,
             rng.random((300, 3)) < 0.9]
    for m in cases:
        want = oracle.edt_sq(m)
        got = mz.isotropic.distance_transform_edt_sq(m)
        assert got.dtype == np.int32 and np.array_equal(got, want), m.shape


# ------------------------------------------------------------------------------------------------
def test_label_golden(mz, golden_labels):
    z = golden_labels
    for name in iso_cases(z):
        m = unpack(z, name)
        lab, n = mz.measure.label(m, return_num=True)
        assert lab.dtype == np.int32 and n == int(z[f"{name}/n"])
        assert np.array_equal(lab, z[f"{name}/labels"]), name


@pytest.mark.parametrize("shape,p", [((64, 64), 0.5), ((257, 1023), 0.55), ((1024, 1024), 0.6), ((333, 77), 0.3),
                                     ((2048, 2048), 0.45), ((31, 2000), 0.7), ((2000, 31), 0.7), ((5, 32), 1.0),
                                     ((9, 64), 0.0), ((100, 33), 0.59)])
def test_label_random_vs_oracle(mz, shape, p):
    rng = np.random.default_rng(hash(shape) % 1000)
    m = rng.random(shape) < p
    want, n = oracle.label(m)
    got, gn = mz.measure.label(m, return_num=True)
    assert gn == n
    assert np.array_equal(got, want)


def test_label_batched_blobs_and_dense_frame(mz):
    imgs = mz.synth.synth_batch(9, 24, lo=64, hi=300)
    imgs.append(mz.synth.synth_dense_frame(4, size=1536, n_blobs=900))
    geom = mz.device.BatchGeometry.from_images(imgs)
    batch = mz.device.DeviceBatch(geom)
    d = batch.upload(geom.pack_host(imgs))
    bits, flags = batch.threshold_pack(d, 40)
    labels, lab_off = batch.label(bits)
    lab = labels.cpu().numpy()
    off = lab_off.cpu().numpy()
    for i, im in enumerate(imgs):
        want, n = oracle.label(im > 40)
        assert off[i + 1] - off[i] == n
        assert np.array_equal(geom.view(lab, i), want), i
    assert off[-1] - off[-2] > 500  # the dense frame has hundreds of labels


def test_label_filters(mz):
    imgs = mz.synth.synth_batch(19, 6, lo=60, hi=200)
    for im in imgs:
        lab, _ = oracle.label(im > 30)
        for min_size in (1, 5, 50):
            want = oracle.remove_small_objects(lab.copy(), min_size)
            work = lab.copy()
            ret = mz.measure.remove_small_objects(work, min_size=min_size, out=work)
            assert ret is work and np.array_equal(work, want)
        want = oracle.clear_border(lab.copy())
        work = lab.copy()
        ret = mz.measure.clear_border(work, out=work)
        assert ret is work and np.array_equal(work, want)


# ------------------------------------------------------------------------------------------------
def test_merge_labels_golden(mz, golden_merge):
    z, meta = golden_merge
    checked = raised = 0
    for key, info in meta.items():
        name, kw = merge_case_args(key)
        lab = z[f"{name}/input"]
        work = lab.copy()
        if "index" in kw:
            idx = [int(v) for v in kw["index"].split("-")]
            call = lambda: mz.merge_labels(work, index=list(idx), max_distance=20, return_merge_distances=True,
                                           labels_out=work)
        else:
            md = None if kw["md"] == "None" else float(kw["md"])
            alias = kw["alias"] == "1"
            call = lambda: mz.merge_labels(work, max_distance=md, path_tolerance=float(kw["tol"]),
                                           return_merge_distances=True, labels_out=work if alias else None)
        if info["raises"]:
            with pytest.raises(TypeError):
                call()
            raised += 1
            continue
        res, dists = call()
        assert np.array_equal(np.asarray(res), z[key + "/labels"]), key
        assert np.array_equal(np.asarray(dists, np.float64), z[key + "/dists"]), key
        if "identity" in info:
            assert (res is work) == info["identity"], key
        checked += 1
    assert checked > 300 and raised == 5


def test_merge_labels_index_list_is_popped_like_the_reference(mz, golden_merge):
    z, _ = golden_merge
    lab = z["sparse/input"]
    for idx in ([5, 1, 3, 8], [8, 7, 6, 5, 4, 3, 2, 1]):
        a, b = list(idx), list(idx)
        scipy_chain.merge_labels(lab.copy(), index=a, max_distance=20)
        mz.merge_labels(lab.copy(), index=b, max_distance=20)
        assert a == b


def test_merge_labels_batched_vs_oracle(mz):
    imgs = mz.synth.synth_batch(23, 12, lo=48, hi=160)
    labs = []
    for im in imgs:
        m = oracle.isotropic_closing(oracle.isotropic_opening(im > 40, 1), 2)
        labs.append(oracle.label(m)[0])
    geom = mz.device.BatchGeometry.from_images(labs)
    batch = mz.device.DeviceBatch(geom)
    d_lab = batch.upload(geom.pack_host(labs, dtype=np.int32))
    bounds = [int(l.max()) for l in labs]
    lab_off, n_obj = batch.lab_off_from_bounds(bounds)
    md, nm, _, status, _ = batch.merge_labels(d_lab, d_lab, lab_off, n_obj, 10.0)
    got = d_lab.cpu().numpy()
    md, nm, off = md.cpu().numpy(), nm.cpu().numpy(), lab_off.cpu().numpy()
    for i, l in enumerate(labs):
        work = l.copy()
        want, dists = oracle.merge_labels(work, max_distance=10.0, return_merge_distances=True, labels_out=work)
        assert np.array_equal(geom.view(got, i), want), i
        assert list(md[off[i]:off[i] + nm[i]]) == list(dists)
    assert (status.cpu().numpy() == 0).all()


def _merge_stress_images(rng):
    """Label images that stress the windowed merge kernel: sparse specks, touching mosaics (bridges overwrite other
    labels: boxes and minima must be recomputed), more labels than the shared-memory tables hold, a label that fills
    the frame."""
    out = []
    for k in range(6):  # specks around a blob
        h, w = int(rng.integers(40, 160)), int(rng.integers(40, 160))
        m = rng.random((h, w)) < 0.012
        yy, xx = np.mgrid[:h, :w]
        m |= (yy - h / 2) ** 2 / (h / 5) ** 2 + (xx - w / 2) ** 2 / (w / 4) ** 2 < 1
        out.append(oracle.label(m)[0])
    for k in range(6):  # mosaics: labels touch, every label has several pieces
        h, w = int(rng.integers(24, 90)), int(rng.integers(24, 90))
        cells = rng.integers(0, 9, (h // 4 + 1, w // 4 + 1))
        lab = np.kron(cells, np.ones((4, 4), np.int64))[:h, :w]
        lab[rng.random((h, w)) < 0.5] = 0
        out.append(lab.astype(np.int32))
    lab = np.zeros((200, 250), np.int32)  # 1100 single pixels on a grid: tables in global memory
    ys, xs = np.mgrid[2:200:6, 2:250:7]
    ys, xs = ys.ravel()[:1100], xs.ravel()[:1100]
    lab[ys, xs] = rng.permutation(len(ys)) + 1
    out.append(lab)
    lab = np.ones((50, 70), np.int32)  # one label fills the frame, others inside it
    lab[10:14, 10:14] = 2
    lab[10:14, 20:24] = 0
    lab[11:13, 21:23] = 3
    out.append(lab)
    lab = np.zeros((180, 40), np.int32)  # tall windows (several 32-row segments), gaps inside columns
    lab[5:60:9, 5:30] = 1
    lab[66:170:13, 8:12] = 2
    lab[100:175, 18:21] = 3
    out.append(lab)
    h, w = 260, 300  # large windows: the lower-envelope row pass; an irregular blob with holes, satellites around it
    yy, xx = np.mgrid[:h, :w]
    m = np.zeros((h, w), bool)
    for cy, cx, ry, rx in ((130, 150, 80, 60), (90, 90, 30, 70), (190, 200, 40, 50), (60, 230, 25, 25)):
        m |= (yy - cy) ** 2 / ry ** 2 + (xx - cx) ** 2 / rx ** 2 < 1
    m &= ~((yy - 140) ** 2 + (xx - 150) ** 2 < 400)
    m |= rng.random((h, w)) < 0.004
    out.append(oracle.label(m)[0])
    lab = np.zeros((215, 20), np.int32)  # ONE bridge (bar 1 - line 2) takes a pixel from each of 35 labels at once:
    lab[:, 0:3] = 1                      # more than the kernel lists -> boxes and minima of all labels are rebuilt
    lab[:, 9] = 2
    for k, r in enumerate(range(3, 213, 6)):
        lab[r, 11:14] = 3 + k
    out.append(lab)
    # a one-pixel label swallowed by a bridge while later labels remain: its minimum is `initial`, it is never popped
    # (a crop of a synthetic bench vignette on which the first version of the kernel raised)
    out.append(np.load(os.path.join(os.path.dirname(__file__), "golden", "merge_swallowed_label.npy")).astype(np.int32))
    return out


@pytest.mark.parametrize("md,tol", [(10.0, 5.0), (3.0, 5.0), (12.5, 0.0), (6.0, 1.5)])
@pytest.mark.parametrize("alias", [True, False])
def test_merge_labels_windowed_stress(mz, md, tol, alias):
    """The windowed kernel (index = None, max_distance given) against the oracle on images built to hit its corner
    cases; `tol >= md + 2` makes the bridge condition hold OUTSIDE the distance window."""
    labs = _merge_stress_images(np.random.default_rng(int(md * 10 + tol)))
    geom = mz.device.BatchGeometry.from_images(labs)
    batch = mz.device.DeviceBatch(geom)
    d_lab = batch.upload(geom.pack_host(labs, dtype=np.int32))
    lab_off, n_obj = batch.lab_off_from_bounds([int(l.max()) for l in labs])
    d_out = d_lab if alias else d_lab.clone()
    dist, nm, _, status, _ = batch.merge_labels(d_lab, d_out, lab_off, n_obj, md, tol)
    got = d_out.cpu().numpy()
    dist, nm, off, status = dist.cpu().numpy(), nm.cpu().numpy(), lab_off.cpu().numpy(), status.cpu().numpy()
    for i, l in enumerate(labs):
        work = l.copy()
        try:
            want, dists = oracle.merge_labels(work, max_distance=md, path_tolerance=tol, return_merge_distances=True,
                                              labels_out=work if alias else None)
        except TypeError:
            assert status[i] != 0, i
            continue
        assert status[i] == 0, i
        assert np.array_equal(geom.view(got, i), want), (i, md, tol, alias)
        assert list(dist[off[i]:off[i] + nm[i]]) == list(dists), i


# ------------------------------------------------------------------------------------------------
def assert_tables_close(got, want, with_image=True):
    F = oracle
    assert got.shape == want.shape
    for g, w in zip(got, want):
        assert g[F.F_LABEL] == w[F.F_LABEL] and g[F.F_AREA] == w[F.F_AREA]
        if w[F.F_AREA] == 0:
            continue
        area = w[F.F_AREA]
        assert np.array_equal(g[F.F_BBOX:F.F_BBOX + 4], w[F.F_BBOX:F.F_BBOX + 4])
        np.testing.assert_allclose(g[F.F_CENTROID:F.F_CENTROID + 2], w[F.F_CENTROID:F.F_CENTROID + 2], rtol=1e-14)
        gm, wm = g[F.F_MU:F.F_MU + 16].reshape(4, 4), w[F.F_MU:F.F_MU + 16].reshape(4, 4)
        gn, wn = g[F.F_NU:F.F_NU + 16].reshape(4, 4), w[F.F_NU:F.F_NU + 16].reshape(4, 4)
        for p in range(4):
            for q in range(4):
                scale = area ** ((p + q) / 2 + 1)
                assert abs(gm[p, q] - wm[p, q]) <= 1e-5 * abs(wm[p, q]) + 1e-9 * scale, (p, q, gm[p, q], wm[p, q])
                if p + q >= 2:
                    assert abs(gn[p, q] - wn[p, q]) <= 1e-5 * abs(wn[p, q]) + 1e-9, (p, q)
                else:
                    assert np.isnan(gn[p, q]) and np.isnan(wn[p, q])
        np.testing.assert_allclose(g[F.F_HU:F.F_HU + 7], w[F.F_HU:F.F_HU + 7], rtol=1e-5, atol=1e-12)
        for c in (F.F_EIG, F.F_EIG + 1, F.F_AXIS_MAJOR, F.F_AXIS_MINOR, F.F_T00, F.F_T01, F.F_T11):
            np.testing.assert_allclose(g[c], w[c], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(g[F.F_ECC], w[F.F_ECC], rtol=1e-5, atol=2e-7)
        if abs(w[F.F_T00] - w[F.F_T11]) > 1e-6 * max(1.0, abs(w[F.F_T00])) or w[F.F_T00] == w[F.F_T11]:
            d = abs(g[F.F_ORIENT] - w[F.F_ORIENT])
            assert min(d, abs(d - np.pi)) <= 1e-5, (g[F.F_ORIENT], w[F.F_ORIENT])
        if with_image:
            assert g[F.F_IMIN] == w[F.F_IMIN] and g[F.F_IMAX] == w[F.F_IMAX]
            np.testing.assert_allclose(g[F.F_IMEAN], w[F.F_IMEAN], rtol=1e-14)
            np.testing.assert_allclose(g[F.F_FRAC_INVALID], w[F.F_FRAC_INVALID], rtol=1e-14)


def test_regionprops_golden_cases(mz, golden_props):
    z, _ = golden_props
    for name in ("ellipse", "square", "pixel", "hline", "blobs"):
        lab, inten = z[f"{name}/labels"], z[f"{name}/intensity"]
        want = oracle.regionprops_table(lab, inten)
        got = mz.measure.regionprops_table(lab, inten)
        assert_tables_close(got, want)
    lab = np.zeros((12, 12), np.int32)
    lab[2:8, 3:9] = 1
    t = mz.measure.regionprops_table(lab)[0]
    assert t[oracle.F_ORIENT] == -np.pi / 4          # T00 == T11 branch, exact
    lab = np.zeros((5, 5), np.int32)
    lab[2, 3] = 1
    t = mz.measure.regionprops_table(lab)[0]
    assert t[oracle.F_ECC] == 0 and t[oracle.F_AXIS_MAJOR] == 0


def test_regionprops_random_label_images_and_gaps(mz):
    rng = np.random.default_rng(11)
    # arbitrary label image: noise labels with gaps (labels 3 and 6 absent), intensity with zeros
    lab = rng.integers(0, 9, size=(70, 101)).astype(np.int32)
    lab[lab == 3] = 0
    lab[lab == 6] = 0
    inten = rng.integers(0, 256, size=lab.shape, dtype=np.uint8)
    inten[rng.random(lab.shape) < 0.2] = 0
    assert_tables_close(mz.measure.regionprops_table(lab, inten), oracle.regionprops_table(lab, inten))
    assert_tables_close(mz.measure.regionprops_table(lab), oracle.regionprops_table(lab), with_image=False)
    # big coordinates: a frame-sized label image
    img = mz.synth.synth_dense_frame(2, size=2048, n_blobs=600)
    lab, n = oracle.label(img > 40)
    assert_tables_close(mz.measure.regionprops_table(lab, img), oracle.regionprops_table(lab, img))


def test_mask_properties_image_properties_semantics(mz):
    img = mz.synth.synth_batch(5, 1, size=(120, 90))[0]
    mask = img > 40
    want = oracle.regionprops_table(mask.astype(np.int32), img)
    assert_tables_close(mz.measure.mask_properties(mask, img), want)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("merge", [0, 10])
@pytest.mark.parametrize("filters", [False, True])
def test_stage_composite_vs_reference_chain(mz, merge, filters):
    S = mz.stage
    imgs = mz.synth.synth_batch(101 + merge, 20, lo=64, hi=320)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, merge_segments_distance=merge,
                                            min_area=12 if filters else 0, clear_border=filters)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
    res = st(imgs)
    assert len(res) == len(imgs)
    for i, im in enumerate(imgs):
        mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2, clear_border_flag=filters,
                                                     min_area=12 if filters else 0, merge_segments_distance=merge)
        assert np.array_equal(res.mask(i), mask), i
        assert np.array_equal(res.labels(i), labels), i
        feats = res.features(i)
        nrow = min(len(feats), len(table))
        assert len(feats) >= len(table)
        assert_tables_close(feats[:nrow], table[:nrow])
        assert (feats[nrow:, oracle.F_AREA] == 0).all()  # labels removed by the filters leave empty rows


def test_stage_threshold_branch(mz):
    S = mz.stage
    imgs = mz.synth.synth_batch(7, 8, lo=64, hi=200)
    imgs.append(np.zeros((50, 60), np.uint8))  # empty mask -> dropped by the Filter (loki/pipeline.py:651)
    st = S.LokiSegmentationStage(threshold=S.ThresholdSegmentationConfig(35.5))
    res = st(imgs)
    assert list(res.keep) == [True] * 8 + [False]
    for i, im in enumerate(imgs[:8]):
        mask = im > 35.5
        assert np.array_equal(res.mask(i), mask)
        assert_tables_close(res.features(i), oracle.regionprops_table(mask.astype(np.int32), im))


def test_stage_postprocess_branch_on_model_prediction(mz):
    S = mz.stage
    imgs = mz.synth.synth_batch(8, 6, lo=64, hi=200)
    preds = [(im > 50).astype(np.float32) * 0.9 for im in imgs]  # "foreground_pred" of a model
    pp = S.SegmentationPostprocessingConfig(closing_radius=3, opening_radius=2)
    res = S.LokiSegmentationStage(postprocess=pp)(imgs, foreground_pred=preds)
    for i, (im, pr) in enumerate(zip(imgs, preds)):
        m = scipy_chain.closing(scipy_chain.opening(np.asarray(pr, dtype=bool), 2), 3)
        lab, _ = scipy_chain.label(m)
        assert np.array_equal(res.mask(i), m) and np.array_equal(res.labels(i), lab)


def test_full_size_properties_without_oracle(mz):
    """Size-independent properties on a BASELINE-sized vignette mix (1024-px sides)."""
    S = mz.stage
    imgs = mz.synth.synth_batch(55, 6, lo=700, hi=1024)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    res = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)(imgs)
    for i in range(len(imgs)):
        lab, mask = res.labels(i), res.mask(i)
        assert np.array_equal(lab > 0, mask)
        flat = lab.ravel()
        first = flat[np.sort(np.unique(flat, return_index=True)[1])]
        first = first[first > 0]
        assert np.array_equal(first, np.arange(1, len(first) + 1))  # raster order of first pixels
        feats = res.features(i)
        assert feats[:, oracle.F_AREA].sum() == mask.sum()            # areas partition the mask
        assert np.array_equal(feats[:, oracle.F_AREA], np.bincount(flat)[1:])


def _edge_images(mz):
    rng = np.random.default_rng(77)
    imgs = mz.synth.synth_batch(202, 10, lo=64, hi=400)
    imgs += [np.zeros((70, 90), np.uint8), np.full((65, 64), 255, np.uint8), np.full((1, 1), 200, np.uint8),
             np.full((1, 77), 99, np.uint8), np.full((130, 1), 99, np.uint8),
             (rng.random((200, 300)) < 0.5).astype(np.uint8) * 200,      # noise: more runs than UF slots -> fallback
             (rng.random((96, 33)) < 0.9).astype(np.uint8) * 200,
             mz.synth.synth_batch(203, 1, size=(1024, 1000))[0],          # too large for the fused kernel
             mz.synth.synth_batch(204, 1, size=(580, 1016))[0],           # largest fused class
             np.full((40, 40), 41, np.uint8)]                             # all ones after threshold (phantom in erosion)
    faint = np.zeros((64, 64), np.uint8)
    faint[10, 10] = 255                                                   # eroded away -> empty plane feeds the dilation
    imgs.append(faint)
    # multi-band vignettes of the band pipeline: uniform planes (scipy's phantom pixel -> flagged, per-operator redo),
    # foreground in a single band, a component crossing every band seam, a tall narrow and a wide flat one
    imgs += [np.full((300, 800), 255, np.uint8), np.zeros((500, 700), np.uint8)]
    one = np.zeros((600, 640), np.uint8)
    one[500:520, 100:300] = 250
    imgs.append(one)
    bar = np.zeros((900, 700), np.uint8)
    bar[5:895, 340:352] = 250
    bar[100:110, 10:690] = 250
    bar[700:703, 10:690] = 250
    imgs.append(bar)
    imgs += [mz.synth.synth_batch(205, 1, size=(3000, 70))[0], mz.synth.synth_batch(206, 1, size=(70, 3000))[0]]
    return imgs


@pytest.mark.parametrize("pipeline", ["bands", "fused"])
@pytest.mark.parametrize("radii", [(1, 2), (2, 3), (0, 2), (3, 0), (0, 0), (1.5, 2.5)])
def test_fused_vignette_kernel_equals_reference_chain(mz, radii, pipeline):
    S = mz.stage
    r_open, r_close = radii
    imgs = _edge_images(mz)
    pp = S.SegmentationPostprocessingConfig(closing_radius=r_close, opening_radius=r_open)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, pipeline=pipeline)
    res = st(imgs)
    for i, im in enumerate(imgs):
        mask, labels, table = scipy_chain.loki_chain(im, 40, r_open, r_close)
        assert np.array_equal(res.mask(i), mask), (i, im.shape)
        assert np.array_equal(res.labels(i), labels), (i, im.shape)
        assert len(res.features(i)) == len(table)
        assert_tables_close(res.features(i), table)


def test_fused_and_generic_paths_agree_bitwise(mz):
    S = mz.stage
    imgs = _edge_images(mz)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, min_area=9, clear_border=True)
    a = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, fused=True)(imgs)
    am = [a.mask(i).copy() for i in range(len(imgs))]
    al = [a.labels(i).copy() for i in range(len(imgs))]
    at, ao = a.table.copy(), a.lab_off.copy()
    b = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, fused=False)(imgs)
    assert np.array_equal(ao, b.lab_off)
    for i in range(len(imgs)):
        assert np.array_equal(am[i], b.mask(i)) and np.array_equal(al[i], b.labels(i))
    assert np.array_equal(np.nan_to_num(at, nan=-1.0)[:, :8], np.nan_to_num(b.table, nan=-1.0)[:, :8])


def test_streaming_map_equals_single_calls(mz):
    S = mz.stage
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
    batches = [mz.synth.synth_batch(300 + b, 5 + b, lo=64, hi=260) for b in range(5)]
    batches.insert(2, [])  # an empty batch in the stream
    got = []
    for res in st.map(batches):
        got.append(([res.mask(i).copy() for i in range(len(res))], [res.labels(i).copy() for i in range(len(res))],
                    res.table.copy(), res.lab_off.copy()))
    assert len(got) == len(batches)
    for imgs, (masks, labels, table, off) in zip(batches, got):
        assert len(masks) == len(imgs)
        rows = 0
        for i, im in enumerate(imgs):
            m, l, t = scipy_chain.loki_chain(im, 40, 1, 2)
            assert np.array_equal(masks[i], m) and np.array_equal(labels[i], l)
            assert off[i + 1] - off[i] == len(t)
            assert_tables_close(table[off[i]:off[i + 1]], t)
            rows += len(t)
        assert rows == len(table)


def test_config3_radius_sweep_on_2048_frame(mz):
    """BASELINE.json configs[2]: isotropic closing / opening on a 2048 x 2048 frame (bit-plane disk path up to
    r = 32, exact-EDT path beyond)."""
    rng = np.random.default_rng(3)
    frame = mz.synth.synth_dense_frame(7, size=2048, n_blobs=60)
    frame = np.maximum(frame, (rng.random(frame.shape) < 0.002).astype(np.uint8) * 255)  # specks for the opening
    mask = frame > 40
    for r in (1, 5, 32, 40):
        for op in ("opening", "closing"):
            want = getattr(oracle, f"isotropic_{op}")(mask, r)
            got = getattr(mz.isotropic, f"isotropic_{op}")(mask, r)
            assert np.array_equal(got, want), (op, r)


def test_config4_dense_4096_frame_label_and_regionprops(mz):
    """BASELINE.json configs[3]: a 4096 x 4096 frame with thousands of labels through the stage (too large for
    the vignette-resident kernel: per-operator kernels) and through the per-image drop-ins."""
    S = mz.stage
    frame = mz.synth.synth_dense_frame(11, size=4096, n_blobs=3000)
    want_lab, n = oracle.label(frame > 40)
    assert n > 2000
    want_tab = oracle.regionprops_table(want_lab, frame)
    pp = S.SegmentationPostprocessingConfig()  # no morphology, no filters
    res = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)([frame])
    assert np.array_equal(res.labels(0), want_lab)
    assert np.array_equal(res.mask(0), frame > 40)
    assert len(res.table) == n
    assert_tables_close(res.table, want_tab)
    lab2, n2 = mz.measure.label(frame > 40, return_num=True)
    assert n2 == n and np.array_equal(lab2, want_lab)
    # the frame takes the band pipeline with the global-memory labelling kernels (no per-operator redo): in compact
    # form it comes down as a run list
    resc = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=True)([frame])
    assert resc.compact and 0 not in resc._dense
    assert np.array_equal(resc.labels(0), want_lab) and len(resc.table) == n
    assert_tables_close(resc.table, want_tab)
    # and with the default morphology on a smaller frame (still labelled in global memory: >= 2 MPix)
    f2 = mz.synth.synth_dense_frame(12, size=2300, n_blobs=900)[:2200, :]
    mask, labels, table = scipy_chain.loki_chain(f2, 40, 1, 2)
    pp2 = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    for compact in (False, True):
        r2 = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp2, compact=compact)([f2, frame[:70, :90]])
        assert np.array_equal(r2.mask(0), mask) and np.array_equal(r2.labels(0), labels)
        assert_tables_close(r2.features(0), table)
        if compact:
            assert 0 not in r2._dense


def test_stream_objects_adapter_and_odd_geometries(mz):
    """The stream adapter (input order, per-object metadata) and vignettes the fused kernel must hand to the
    per-operator path: wider than 65535 px, one-pixel rows / columns."""
    S = mz.stage
    rng = np.random.default_rng(2)
    imgs = mz.synth.synth_batch(77, 5, lo=64, hi=160)
    wide = (rng.random((3, 70000)) < 0.3).astype(np.uint8) * 200
    imgs += [wide, np.full((1, 300), 90, np.uint8), np.full((257, 1), 90, np.uint8)]
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp)
    objs = [{"image": im, "meta": {"object_id": f"v{k}"}, "k": k} for k, im in enumerate(imgs)]
    out = list(S.stream_objects(st, objs, batch_size=3, padding=0))
    assert [o["k"] for o in out] == list(range(len(imgs)))
    for o, im in zip(out, imgs):
        mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2)
        assert np.array_equal(o["mask"], mask) and np.array_equal(o["labels"], labels)
        live = table[table[:, oracle.F_AREA] > 0]
        assert [d["object_sequence"] for d in o["objects"]] == [int(v) for v in live[:, oracle.F_LABEL]]
        assert [d["object_area"] for d in o["objects"]] == [float(v) for v in live[:, oracle.F_AREA]]
        assert all(d["object_id"] == f"v{o['k']}" for d in o["objects"])


def test_config1_thousand_256x256_vignettes(mz):
    """BASELINE.json configs[0]: 1 000 synthetic 256 x 256 uint8 vignettes (seed 0), default stage parameters of
    SURVEY.md 8d, with and without merge_segments_distance; every 16th vignette against the reference chain, all of
    them through size-independent properties."""
    S = mz.stage
    imgs = mz.synth.synth_batch(0, 1000, size=(256, 256))
    for merge in (0, 10):
        pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, merge_segments_distance=merge)
        res = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, merge_errors="ignore")(imgs)
        failed = set() if res.merge_failed is None else set(int(i) for i in res.merge_failed)
        for i in range(0, 1000, 16):
            try:
                mask, labels, table = scipy_chain.loki_chain(imgs[i], 40, 1, 2, merge_segments_distance=merge)
            except TypeError:
                assert i in failed  # the reference raises for this vignette (merge_labels.py:19-20)
                continue
            assert i not in failed
            assert np.array_equal(res.mask(i), mask) and np.array_equal(res.labels(i), labels), (merge, i)
            feats = res.features(i)
            k = len(table)
            assert_tables_close(feats[:k], table)
        total_area = 0
        for i in range(1000):
            lab = res.labels(i)
            if merge == 0:
                assert np.array_equal(lab > 0, res.mask(i))
            total_area += int((lab > 0).sum())
        assert res.table[:, oracle.F_AREA].sum() == total_area


# ---- shape features: perimeter / filled_area / euler_number -------------
def assert_shape_equal(got, want):
    assert got.shape == want.shape
    absent = np.isnan(want[:, 0])
    assert np.array_equal(np.isnan(got[:, 0]), absent)
    g, w = got[~absent], want[~absent]
    assert np.array_equal(g[:, 1:7], w[:, 1:7])          # filled_area, euler_number, n1, n2, n3, convex_area: exact integers
    assert np.allclose(g[:, 0], w[:, 0], rtol=1e-12, atol=0)


def test_label_shape_special_cases(mz):
    from oracle import shape as oshape
    ring = np.ones((7, 7), np.int32)
    ring[2:5, 2:5] = 0
    leak = ring.copy()
    leak[0, 0] = leak[0, 1] = leak[1, 0] = leak[1, 1] = 0    # hole 8-connected to the outside: not filled
    nested = np.zeros((40, 70), np.int32)                     # island (label 2) with its own hole inside a hole of 1
    nested[2:38, 3:60] = 1
    nested[6:34, 8:55] = 0
    nested[10:30, 12:50] = 2
    nested[14:26, 20:40] = 0
    nested[18:22, 25:30] = 5                                  # labels 3, 4 absent
    spiral = np.zeros((33, 33), np.int32)                     # the flood has to wind inwards
    for k in range(0, 16, 2):
        spiral[k, k:33 - k] = 1
        spiral[k:33 - k, 32 - k] = 1
        spiral[32 - k, k:33 - k] = 1
        spiral[k + 2:33 - k, k] = 1
    wide = np.zeros((9, 1500), np.int32)                      # rows of more than 32 words: carries across chunks
    wide[1, 1:1499] = 1
    wide[7, 1:1499] = 1
    wide[1:8, 1] = 1
    wide[1:8, 1498] = 1
    wide[3:6, 700:900] = 2
    for lab in (ring, leak, nested, spiral, wide, np.eye(6, dtype=np.int32), np.ones((1, 1), np.int32),
                np.ones((1, 40), np.int32), np.ones((300, 2), np.int32)):
        assert_shape_equal(mz.measure.regionprops_shape(lab), oshape.label_shape(lab))
    # ImageProperties semantics: the whole mask is one region
    m = nested > 0
    assert_shape_equal(mz.measure.mask_shape(m), oshape.label_shape(m.astype(np.int32)))


def test_label_shape_random_and_large(mz):
    from oracle import shape as oshape
    rng = np.random.default_rng(21)
    for shape, p in (((70, 101), 0.5), ((200, 333), 0.62), ((64, 64), 0.9), ((31, 257), 0.3)):
        lab, _ = oracle.label(rng.random(shape) < p)          # noise near the percolation threshold: many holes
        assert_shape_equal(mz.measure.regionprops_shape(lab), oshape.label_shape(lab))
    lab = rng.integers(0, 6, size=(90, 120)).astype(np.int32)  # arbitrary label image: labels touch each other
    assert_shape_equal(mz.measure.regionprops_shape(lab), oshape.label_shape(lab))
    # crops too large for shared memory (planes in the CTA's slab), thousands of labels
    img = mz.synth.synth_dense_frame(3, size=1536, n_blobs=400)
    lab, _ = oracle.label(img > 40)
    assert_shape_equal(mz.measure.regionprops_shape(lab), oshape.label_shape(lab))
    big = np.zeros((1024, 1024), np.int32)
    big[rng.random(big.shape) < 0.6] = 1                       # one label whose crop is the whole 1024^2 vignette
    assert_shape_equal(mz.measure.regionprops_shape(big), oshape.label_shape(big))


def test_stage_shape_features_and_zooprocess_keys(mz):
    from oracle import shape as oshape
    from maze_image_processing_pipeline_b200.regions import objects_of
    S = mz.stage
    imgs = mz.synth.synth_batch(77, 24, lo=64, hi=400)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, shape_features=True)
    outs = list(st.map([imgs[:12], imgs[12:]]))
    k = 0
    for res in outs:
        for i in range(len(res)):
            _, labels, table = scipy_chain.loki_chain(imgs[k], 40, 1, 2)
            want = oshape.label_shape(labels, max_label=len(res.features(i)))
            assert_shape_equal(res.shape_features(i), want)
            for o, row in zip(objects_of(res, i, image=imgs[k]), want[~np.isnan(want[:, 0])]):
                assert o["object_area"] == row[1] and o["object_perim."] == pytest.approx(row[0], rel=1e-12)
                assert o["object_euler_number"] == row[2] and o["object_area_exc"] <= o["object_area"]
                assert o["object_convex_area"] == row[6] and o["object_solidity"] == o["object_area_exc"] / row[6]
            k += 1
    assert k == 24
    # threshold branch: one region per vignette, taken from the bit plane
    st = S.LokiSegmentationStage(threshold=S.ThresholdSegmentationConfig(35.5), shape_features=True)
    res = st(imgs[:6])
    for i in range(6):
        assert_shape_equal(res.shape_features(i), oshape.label_shape((imgs[i] > 35.5).astype(np.int32), max_label=1))


# ---- footprint morphology: what the live pipeline calls (loki/pipeline.py:408-427) -----------
def test_binary_morphology_with_footprints_vs_scipy(mz):
    from maze_image_processing_pipeline_b200 import morphology as M
    rng = np.random.default_rng(31)
    masks = [rng.random((61, 97)) < 0.55, rng.random((40, 40)) < 0.9, rng.random((1, 70)) < 0.8,
             rng.random((50, 1)) < 0.8, np.ones((20, 33), bool), np.zeros((7, 9), bool),
             mz.synth.synth_batch(3, 1, size=(150, 131))[0] > 40]
    fps = [None, M.disk(1), M.disk(2), M.disk(5), M.disk(9), np.ones((3, 3), np.uint8), np.ones((3, 7), np.uint8),
           M.disk(1, decomposition="crosses"), M.disk(3, decomposition="crosses"), M.disk(13, decomposition="crosses")]
    default = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)
    for m in masks:
        for fp in fps:
            ofp = default if fp is None else fp
            for op in ("erosion", "dilation", "opening", "closing"):
                want = getattr(scipy_chain, f"binary_{op}")(m, ofp)
                got = getattr(M, f"binary_{op}")(m, fp)
                assert got.dtype == bool and np.array_equal(got, want), (m.shape, op)
    buf = np.zeros(masks[0].shape, bool)
    assert M.binary_opening(masks[0].astype(np.uint8) * 255, M.disk(2), out=buf) is buf
    assert np.array_equal(buf, scipy_chain.binary_opening(masks[0], M.disk(2)))


@pytest.mark.parametrize("fused", [True, False])
def test_stage_crosses_mode_vs_live_pipeline_chain(mz, fused):
    """LokiSegmentationStage(morphology="crosses"): threshold -> binary_opening(disk(r, "crosses")) ->
    binary_closing(disk(r, "crosses")) -> label, in the fused kernel and through the per-operator kernels."""
    from maze_image_processing_pipeline_b200 import morphology as M
    S = mz.stage
    imgs = mz.synth.synth_batch(55, 14, lo=64, hi=300)
    imgs.append(np.full((70, 90), 255, np.uint8))   # all foreground: no phantom pixel in this mode
    for r_open, r_close in ((1, 2), (3, 13), (0, 4)):
        pp = S.SegmentationPostprocessingConfig(closing_radius=r_close, opening_radius=r_open)
        st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, morphology="crosses", fused=fused)
        res = st(imgs)
        for i, im in enumerate(imgs):
            m = im > 40
            if r_open:
                m = scipy_chain.binary_opening(m, M.disk(r_open, decomposition="crosses"))
            if r_close:
                m = scipy_chain.binary_closing(m, M.disk(r_close, decomposition="crosses"))
            lab, _ = scipy_chain.label(m)
            assert np.array_equal(res.mask(i), m), (r_open, r_close, i)
            assert np.array_equal(res.labels(i), lab), (r_open, r_close, i)
            assert_tables_close(res.features(i), oracle.regionprops_table(lab, im, max_label=len(res.features(i))))
