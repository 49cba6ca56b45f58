"""The compact result transport (run lists instead of dense label images across PCIe), the streaming call on the
paths that share one device workspace, and concurrent callers of the C-ABI."""
import threading

import numpy as np
import pytest

import oracle
from oracle import scipy_chain
from test_gpu_parity import _edge_images, assert_tables_close, mz  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("radii", [(1, 2), (0, 0), (2, 3)])
def test_compact_results_equal_the_reference_chain(mz, radii):
    """LokiSegmentationStage(compact=True): masks / label images expanded on the host from the run list, crops of
    single objects, the whole-batch materialisation and the object table, against the reference chain -- on the edge
    cases too (uniform planes, noise that overflows the run tables, oversize vignettes: those come down dense)."""
    S = mz.stage
    r_open, r_close = radii
    imgs = _edge_images(mz)
    pp = S.SegmentationPostprocessingConfig(closing_radius=r_close, opening_radius=r_open)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=True)
    res = st(imgs)
    assert res.compact
    want = [scipy_chain.loki_chain(im, 40, r_open, r_close) for im in imgs]
    for i, (mask, labels, table) in enumerate(want):
        assert np.array_equal(res.mask(i), mask), (i, imgs[i].shape)
        assert np.array_equal(res.labels(i), labels), (i, imgs[i].shape)
        assert len(res.features(i)) == len(table)
        assert_tables_close(res.features(i), table)
        h, w = labels.shape
        sl = (slice(h // 4, h // 2 + 3), slice(max(0, w // 3 - 2), w))
        assert np.array_equal(res.object_mask(i, sl), mask[sl])
        for l in range(1, min(int(labels.max()), 4) + 1):
            assert np.array_equal(res.object_mask(i, sl, l), labels[sl] == l)
    dense = res.materialize()
    for i, (mask, labels, _) in enumerate(want):
        assert np.array_equal(dense.mask(i), mask) and np.array_equal(dense.labels(i), labels)


def test_compact_streaming_and_objects(mz):
    """stage.map in compact mode (three batches in flight over rotating lanes), consumed through FindRegions /
    recalc_metadata / ZooProcess features with the default padding of 75."""
    S = mz.stage
    from maze_image_processing_pipeline_b200.regions import extract_roi, find_regions, objects_of
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=True)
    batches = [mz.synth.synth_batch(500 + b, 6 + b, lo=64, hi=300) for b in range(5)]
    n = 0
    for imgs, res in zip(batches, st.map(batches)):
        assert res.compact and len(res) == len(imgs)
        for i, im in enumerate(imgs):
            mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2)
            assert np.array_equal(res.labels(i), labels) and np.array_equal(res.mask(i), mask)
            regs = list(find_regions(res, i, padding=75, image=im))
            assert [r.label for r in regs] == [int(v) for v in table[table[:, oracle.F_AREA] > 0][:, oracle.F_LABEL]]
            for r in regs:
                assert np.array_equal(r.image, labels[r.slice] == r.label)
                # ExtractROI (loki/pipeline.py:596-602): the padded crop, masked with the object expanded from its runs
                assert np.array_equal(extract_roi(im, r), im[r.slice])
                assert np.array_equal(extract_roi(im, r, alpha=1, bg_color=7), np.where(labels[r.slice] == r.label, im[r.slice], 7))
                # keep_background: only other objects are painted; the label crop comes from the run list
                assert np.array_equal(r.label_image, labels[r.slice])
                lab = labels[r.slice]
                assert np.array_equal(extract_roi(im, r, alpha=1, bg_color=7, keep_background=True),
                                      np.where((lab == 0) | (lab == r.label), im[r.slice], 7))
            objs = objects_of(res, i, padding=75, image=im)
            assert [o["object_area_exc"] for o in objs] == [float(r.area) for r in regs]
            n += len(objs)
    assert n > 20


@pytest.mark.parametrize("cfg", [dict(clear_border=True, min_area=12), dict(merge_segments_distance=10),
                                 dict(shape=True), dict(threshold_only=True)])
def test_streaming_on_the_shared_workspace_paths(mz, cfg):
    """map() / stream_objects with label filters, merge_labels, shape features and the threshold branch on
    equal-sized consecutive batches: these paths work in ONE device workspace, so a batch must be complete before
    the next one is enqueued (a batch used to overwrite the buffers its predecessor was still downloaded from)."""
    S = mz.stage
    from oracle import shape as oshape
    sizes = [(96, 128), (200, 150), (64, 64), (150, 260)]
    batches = [[mz.synth.synth_batch(900 + 10 * b + k, 1, size=s)[0] for k, s in enumerate(sizes)] for b in range(4)]
    shape = cfg.get("shape", False)
    if cfg.get("threshold_only"):
        st = S.LokiSegmentationStage(threshold=S.ThresholdSegmentationConfig(40))
    else:
        pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1,
                                                clear_border=cfg.get("clear_border", False), min_area=cfg.get("min_area", 0),
                                                merge_segments_distance=cfg.get("merge_segments_distance", 0))
        st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, shape_features=shape, merge_errors="ignore")
    for imgs, res in zip(batches, st.map(batches)):
        failed = set() if res.merge_failed is None else set(int(i) for i in res.merge_failed)
        for i, im in enumerate(imgs):
            if cfg.get("threshold_only"):
                assert np.array_equal(res.mask(i), im > 40)
                assert_tables_close(res.features(i), oracle.regionprops_table((im > 40).astype(np.int32), im))
                continue
            if i in failed:
                continue
            mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2, clear_border_flag=cfg.get("clear_border", False),
                                                         min_area=cfg.get("min_area", 0),
                                                         merge_segments_distance=cfg.get("merge_segments_distance", 0))
            assert np.array_equal(res.mask(i), mask) and np.array_equal(res.labels(i), labels), (i, cfg)
            k = min(len(table), len(res.features(i)))
            assert_tables_close(res.features(i)[:k], table[:k])
            if shape:
                want = oshape.label_shape(labels, max_label=len(res.features(i)))
                got = res.shape_features(i)
                sel = ~np.isnan(want[:, 0])
                assert np.array_equal(got[sel][:, 1:6], want[sel][:, 1:6])


def test_stream_objects_threshold_branch_yields_one_object_per_kept_vignette(mz):
    """loki/pipeline.py:648-656: mask = image > t, vignettes with an empty mask are dropped, ImageProperties makes the
    whole mask ONE region whose ZooProcess features are attached."""
    S = mz.stage
    imgs = mz.synth.synth_batch(61, 7, lo=64, hi=180)
    imgs.insert(3, np.zeros((50, 60), np.uint8))
    st = S.LokiSegmentationStage(threshold=S.ThresholdSegmentationConfig(35.5))
    out = list(S.stream_objects(st, [{"image": im, "meta": {"k": k}, "k": k} for k, im in enumerate(imgs)], batch_size=4))
    assert [o["k"] for o in out] == [0, 1, 2, 4, 5, 6, 7]  # the empty vignette is gone
    for o in out:
        im = imgs[o["k"]]
        mask = im > 35.5
        assert np.array_equal(o["mask"], mask) and o["labels"] is None
        assert len(o["objects"]) == 1
        d = o["objects"][0]
        assert d["object_area_exc"] == mask.sum() and d["k"] == o["k"]
        assert d["object_width"] == im.shape[1] and d["object_height"] == im.shape[0]  # ImageProperties: bbox = frame
    with pytest.raises(TypeError):
        list(S.stream_objects(st, [{"image": imgs[0].astype(np.int32)}]))


def test_two_threads_call_the_stage_concurrently(mz):
    """SURVEY 8b "Threading": the reference may invoke the callables from ThreadPoolExecutor threads
    (loki/pipeline.py:400-402).  Two host threads, each with its own stage object (own streams and workspaces), run
    different batches through the C-ABI at the same time; both must get exactly the single-threaded results."""
    S = mz.stage
    torch = mz.torch
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    jobs = [[mz.synth.synth_batch(700 + 7 * t + r, 12, lo=64, hi=400) for r in range(6)] for t in range(2)]
    want = [[[scipy_chain.loki_chain(im, 40, 1, 2) for im in imgs] for imgs in job] for job in jobs]
    errors = []

    def worker(t):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=bool(t))
                for rep in range(3):
                    for imgs, exp in zip(jobs[t], want[t]):
                        res = st(imgs)
                        for i, (mask, labels, table) in enumerate(exp):
                            assert np.array_equal(res.mask(i), mask) and np.array_equal(res.labels(i), labels)
                            assert_tables_close(res.features(i), table)
        except BaseException as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def _chain(im):
    return scipy_chain.loki_chain(im, 40, 1, 2)


def test_large_vignettes_of_the_bench_distribution_against_the_oracle(mz):
    """configs[1] sizes: 40 vignettes drawn from the benchmark's OWN size distribution (seed 1, log-uniform 64-1024)
    among those with more than 9216 bit-plane words -- the multi-band vignettes of the band pipeline, the three
    largest size classes of the vignette-resident kernel -- plus the largest ones of the job (1024 x 1024 and
    neighbours: the oversize path of the vignette-resident kernel), bit for bit against the reference chain, through
    both pipelines and both transports.  The oracle runs on a process pool (scipy's EDT takes ~1 s per megapixel)."""
    import multiprocessing as mp
    import bench
    S = mz.stage
    hs, ws = bench.job_sizes()
    words = hs.astype(np.int64) * ((ws.astype(np.int64) + 31) // 32)
    big = np.nonzero(words > 9216)[0]
    rng = np.random.default_rng(11)
    pick = list(rng.choice(big, 34, replace=False)) + list(np.argsort(-words)[:6])
    imgs = [mz.synth.synth_vignette(np.random.default_rng(1000 + int(i)), int(hs[i]), int(ws[i])) for i in pick]
    assert sum(im.size > 906_000 for im in imgs) >= 4  # beyond the 28 320-word class of the vignette-resident kernel
    with mp.get_context("fork").Pool(min(16, len(imgs))) as pool:
        want = pool.map(_chain, imgs, chunksize=1)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    for kw in (dict(pipeline="bands"), dict(pipeline="bands", compact=True), dict(pipeline="fused")):
        res = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, **kw)(imgs)
        for i, (mask, labels, table) in enumerate(want):
            assert np.array_equal(res.mask(i), mask), (kw, i, imgs[i].shape)
            assert np.array_equal(res.labels(i), labels), (kw, i, imgs[i].shape)
            assert len(res.features(i)) == len(table)
            assert_tables_close(res.features(i), table)


def _iso(args):
    op, mask, r = args
    return getattr(scipy_chain, op)(mask, r)


def test_full_radius_sweep_1_to_32_and_beyond(mz):
    """BASELINE.json configs[2]: isotropic opening AND closing for EVERY radius 1..32 (plus half-integer and large
    radii up to 120) against the reference's EDT compare -- the bit-plane disk kernel for small radii, the separable
    vertical-distance / row-test pair (maze_morph_pass_wide) from radius 6 on -- on a frame with blobs, specks and
    holes, on uniform planes (scipy's phantom pixel) and on degenerate shapes."""
    import multiprocessing as mp
    rng = np.random.default_rng(9)
    frame = mz.synth.synth_dense_frame(5, size=700, n_blobs=25)[:600, :]
    frame = np.maximum(frame, (rng.random(frame.shape) < 0.003).astype(np.uint8) * 255)
    frame[rng.random(frame.shape) < 0.002] = 0
    masks = [frame > 40, np.ones((90, 140), bool), np.zeros((70, 65), bool), rng.random((1, 300)) < 0.7,
             rng.random((260, 1)) < 0.7, rng.random((130, 97)) < 0.98]
    radii = list(range(1, 33)) + [6.5, 12.5, 40, 64, 120]  # (the separable pair takes over at radius 22)
    jobs = [(op, m, r) for m in masks for r in radii for op in ("opening", "closing")]
    with mp.get_context("fork").Pool(16) as pool:
        want = pool.map(_iso, jobs, chunksize=4)
    for (op, m, r), w in zip(jobs, want):
        got = getattr(mz.isotropic, f"isotropic_{op}")(m, r)
        assert np.array_equal(got, w), (op, m.shape, r)


def test_wide_pass_equals_the_bit_plane_pass_on_footprints_and_batches(mz):
    """maze_morph_pass_wide and maze_morph_pass agree bit for bit (planes and flags) on a packed batch, for disks and
    for registered footprints (the live pipeline's disk(r, "crosses"))."""
    from maze_image_processing_pipeline_b200 import morphology as M
    dev = mz.device
    rng = np.random.default_rng(4)
    imgs = [(rng.random((h, w)) < p).astype(np.uint8) * 255 for h, w, p in
            [(80, 300, 0.97), (257, 33, 0.9), (64, 64, 1.0), (50, 70, 0.0), (1, 1, 1.0), (300, 520, 0.995)]]
    g = dev.BatchGeometry.from_images(imgs)
    b = dev.DeviceBatch(g)
    bits, flags = b.threshold_pack(b.upload(g.pack_host(imgs)), 40)
    codes = [1, 4, 9, 50, 100, 1024] + [M.footprint_pass_code(M.disk(r, decomposition="crosses")) for r in (2, 7, 13)]
    for t in codes:
        for inv in (0, 1):
            vig, n, tiles, nt = b._geo()
            o1, f1 = b.empty_plane(), b.empty_flags()
            dev.check(dev.lib().maze_morph_pass(bits.data_ptr(), o1.data_ptr(), vig, n, tiles, nt, int(t), inv,
                                                flags.data_ptr(), f1.data_ptr(), 0), "maze_morph_pass")
            o2, f2 = b.morph_pass_wide(bits, flags, t, inv)
            mz.torch.cuda.synchronize()
            assert mz.torch.equal(o1, o2), (t, inv)
            assert mz.torch.equal(f1, f2), (t, inv)


@pytest.mark.parametrize("compact", [False, True])
def test_label_filters_on_the_run_list(mz, compact):
    """clear_border / remove_small_objects (loki/pipeline.py:435-448) applied by the labelling kernel on the run list:
    removed labels keep their mask pixels, lose their label (no renumbering) and leave an empty table row -- dense and
    compact transport, on the edge images too (border-touching blobs, specks below min_area, uniform planes)."""
    S = mz.stage
    imgs = _edge_images(mz)[:16] + mz.synth.synth_batch(321, 24, lo=64, hi=500)
    for kw in (dict(clear_border=True), dict(min_area=40), dict(clear_border=True, min_area=12)):
        pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, **kw)
        st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=compact)
        res = st(imgs)
        assert res.compact == compact
        removed = 0
        for i, im in enumerate(imgs):
            mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2, clear_border_flag=kw.get("clear_border", False),
                                                         min_area=kw.get("min_area", 0))
            assert np.array_equal(res.mask(i), mask), (kw, i, im.shape)
            assert np.array_equal(res.labels(i), labels), (kw, i, im.shape)
            f = res.features(i)
            k = min(len(f), len(table))
            assert_tables_close(f[:k], table[:k])
            assert (f[k:, oracle.F_AREA] == 0).all()
            removed += int((f[:, oracle.F_AREA] == 0).sum())
        assert removed > 0, kw


@pytest.mark.parametrize("compact", [False, True])
def test_threshold_branch_on_the_band_pipeline(mz, compact):
    """The branch the reference ships (loki/pipeline.py:648-656): mask = image > t, the whole mask ONE region
    (ImageProperties), empty masks dropped.  Runs on the band pipeline (threshold, run list, single-region
    accumulators), dense or compact; multi-band vignettes, uniform planes and float thresholds included."""
    S = mz.stage
    imgs = _edge_images(mz)
    for thr in (40, 35.5, 254.5, -1):
        st = S.LokiSegmentationStage(threshold=S.ThresholdSegmentationConfig(thr), compact=compact)
        res = st(imgs)
        assert res.compact == compact
        for i, im in enumerate(imgs):
            mask = im > thr
            assert np.array_equal(res.mask(i), mask), (thr, i, im.shape)
            assert res.labels(i) is None
            assert bool(res.keep[i]) == bool(mask.any())
            f = res.features(i)
            assert len(f) == 1
            if mask.any():
                assert_tables_close(f, oracle.regionprops_table(mask.astype(np.int32), im))
            else:
                assert f[0, oracle.F_AREA] == 0


@pytest.mark.parametrize("compact", [False, True])
def test_small_steps_replayed_as_cuda_graphs_follow_their_inputs(mz, compact):
    """Steps of a few vignettes are captured once per lane and argument block and replayed as a CUDA graph
    (maze_stage_step_graph); the plan of the step is cached per lane.  A replay must see the CURRENT contents of the
    input buffer, and another batch must not reuse the cached plan."""
    import torch
    S, device = mz.stage, mz.device
    a = mz.synth.synth_batch(77, 3, lo=90, hi=180)
    b = [np.ascontiguousarray(im[::-1, ::-1]) for im in a]            # same shapes, other contents
    c = mz.synth.synth_batch(78, 4, lo=60, hi=120)                    # another batch on the same lanes
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=compact, n_lanes=2)
    want = {id(x): [scipy_chain.loki_chain(im, 40, 1, 2) for im in x] for x in (a, b, c)}
    geom = device.BatchGeometry.from_images(a)
    batch = st.prepare(device.DeviceBatch(geom))
    d = batch.upload(geom.pack_host(a))
    geom_c = device.BatchGeometry.from_images(c)
    batch_c = st.prepare(device.DeviceBatch(geom_c))
    d_c = batch_c.upload(geom_c.pack_host(c))

    def check(res, g, imgs):
        res.finalize()
        torch.cuda.synchronize()
        labels = None if res.labels is None else res.labels.cpu().numpy()
        table = res.table.cpu().numpy()
        off = res.lab_off.cpu().numpy()
        for i, (mask, lab, tab) in enumerate(want[id(imgs)]):
            if labels is not None and not compact:
                assert np.array_equal(g.view(labels, i), lab), i
            assert off[i + 1] - off[i] == len(tab)
            assert_tables_close(table[off[i]:off[i + 1]], tab)

    for rep in range(8):      # two lanes: calls 0-1 run plainly, 2-3 are captured, 4-7 replay
        imgs = a if rep % 3 != 1 else b
        d.copy_(torch.from_numpy(geom.pack_host(imgs)).to(d.device))
        check(st.run_device(batch, d), geom, imgs)
        if rep == 5:          # another batch in between: its own plan, its own graph
            check(st.run_device(batch_c, d_c), geom_c, c)
    handles = [e for ws in st._ws_ring for e in getattr(ws, "graphs", {}).values()]
    assert st.graphs and any(e[1].value for e in handles), "no step was captured"


@pytest.mark.parametrize("compact", [False, True])
def test_stream_objects_emits_the_extract_roi_crops(mz, compact):
    """stream_objects(rois=True): per object the padded crop and its mask (loki/pipeline.py:596-602), with the
    schema's masking options, from dense and from run-list results."""
    S = mz.stage
    imgs = mz.synth.synth_batch(91, 6, lo=70, hi=160)
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=compact)
    stream = [{"image": im, "meta": {"k": k}, "k": k} for k, im in enumerate(imgs)]
    n = 0
    for o in S.stream_objects(st, stream, batch_size=4, padding=10, rois=True, apply_mask=True, background_color=9,
                              keep_background=True, keep_arrays=False):
        im = imgs[o["k"]]
        _, labels, table = scipy_chain.loki_chain(im, 40, 1, 2)
        live = [int(r[oracle.F_LABEL]) for r in table if r[oracle.F_AREA] > 0]
        assert len(o["rois"]) == len(o["objects"]) == len(live)
        for (crop, mask), meta, l in zip(o["rois"], o["objects"], live):
            ys, xs = np.nonzero(labels == l)
            sl = (slice(max(0, ys.min() - 10), ys.max() + 1 + 10), slice(max(0, xs.min() - 10), xs.max() + 1 + 10))
            lab = labels[sl]
            assert np.array_equal(mask, lab == l)
            assert np.array_equal(crop, np.where((lab == 0) | (lab == l), im[sl], 9))
            assert meta["object_posx"] == sl[1].start and meta["object_posy"] == sl[0].start
            n += 1
    assert n > 6


def test_merge_path_keeps_the_compact_transport(mz):
    """merge_segments_distance > 0 with compact=True: vignettes in which labels merged bring their label images down
    dense (their masks still come from the run list), all others stay run lists; masks, label images, object crops,
    the whole-batch materialisation and the table against the reference chain."""
    S = mz.stage
    from maze_image_processing_pipeline_b200.regions import find_regions
    batches = [mz.synth.synth_batch(300 + b, 24, lo=64, hi=220) for b in range(2)]
    pp = S.SegmentationPostprocessingConfig(closing_radius=2, opening_radius=1, merge_segments_distance=10)
    st = S.LokiSegmentationStage(S.ThresholdSegmentationConfig(40), pp, compact=True, merge_errors="ignore")
    n_dense = n_runs = 0
    for imgs, res in zip(batches, st.map(batches)):
        assert res.compact
        failed = set() if res.merge_failed is None else set(int(i) for i in res.merge_failed)
        want = {}
        for i, im in enumerate(imgs):
            if i in failed:
                continue
            mask, labels, table = scipy_chain.loki_chain(im, 40, 1, 2, merge_segments_distance=10)
            want[i] = (mask, labels)
            assert np.array_equal(res.mask(i), mask) and np.array_equal(res.labels(i), labels), i
            k = min(len(table), len(res.features(i)))
            assert_tables_close(res.features(i)[:k], table[:k])
            for r in find_regions(res, i, padding=5, image=im):
                assert np.array_equal(r.image, labels[r.slice] == r.label)
                assert np.array_equal(r.label_image, labels[r.slice])
            if i in res._dense:
                n_dense += 1
                with pytest.raises(ValueError):
                    res.runs(i)
            else:
                n_runs += 1
                assert len(res.runs(i)) > 0 or not mask.any()
        dense = res.materialize()
        for i, (mask, labels) in want.items():
            assert np.array_equal(dense.mask(i), mask) and np.array_equal(dense.labels(i), labels)
    assert n_dense > 0 and n_runs > n_dense
