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
    from maze_image_processing_pipeline_b200.regions import find_regions, objects_of
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
