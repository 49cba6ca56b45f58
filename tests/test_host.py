"""Host-side logic and the C-ABI surface (CPU only; no kernel is launched)."""
import ctypes
import io
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_geometry_layout():
    from maze_image_processing_pipeline_b200.device import BatchGeometry, TILE_DTYPE, VIG_DTYPE
    g = BatchGeometry([3, 64, 1], [5, 70, 1])
    assert VIG_DTYPE.itemsize == 32 and TILE_DTYPE.itemsize == 8
    assert list(g.vig["wpr"]) == [1, 3, 1]
    assert list(g.pix_off[:3]) == [0, 16, 16 + 4480]
    assert (g.pix_off % 16 == 0).all()
    assert list(g.word_off) == [0, 3, 3 + 192, 3 + 192 + 1]
    assert g.n_tiles == 3 and list(g.tiles["img"]) == [0, 1, 2] and list(g.tiles["word0"]) == [0, 0, 0]
    g2 = BatchGeometry([100], [1000])  # 100 rows x 32 words = 3200 words = 13 tiles
    assert g2.n_tiles == 13 and list(g2.tiles["word0"][:3]) == [0, 256, 512]
    imgs = [np.arange(15, dtype=np.uint8).reshape(3, 5), np.ones((64, 70), np.uint8), np.full((1, 1), 7, np.uint8)]
    flat = g.pack_host(imgs)
    for i, im in enumerate(imgs):
        assert np.array_equal(g.view(flat, i), im)
    with pytest.raises(ValueError):
        BatchGeometry([0], [4])


def test_radius_folding_matches_float64_compare():
    from maze_image_processing_pipeline_b200.device import fold_dilation_radius, fold_erosion_radius, fold_threshold
    ks = np.arange(0, 5000)
    d = np.sqrt(ks.astype(np.float64))
    for r in [0, 0.5, 1, 1.5, 2, 2.5, 3, 5, 8, 13, 32, math.sqrt(2), math.sqrt(5), 2.2360679, 2.23606798, 7.07, 70.7]:
        te, td = fold_erosion_radius(r), fold_dilation_radius(r)
        assert np.array_equal(d > r, ks > te), r
        assert np.array_equal(d < r, ks <= td), r
    assert fold_erosion_radius(-1) == -1 and fold_dilation_radius(0) == -1 and fold_dilation_radius(-2) == -1
    px = np.arange(256)
    for t in (-5, -0.1, 0, 0.5, 30, 30.5, 254.99, 255, 1e9):
        assert np.array_equal(px > t, px > fold_threshold(t)), t


def test_shard_bounds_cover_everything():
    from maze_image_processing_pipeline_b200.stage import shard_bounds
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _header_functions():
    src = open(os.path.join(ROOT, "include", "maze_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(maze_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from maze_image_processing_pipeline_b200 import _lib
    _lib.build()
    handle = ctypes.CDLL(_lib.SO_PATH)
    declared = _header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/maze_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    lib = _lib.lib()
    assert lib.maze_version() >= 100


def test_sass_is_sm100a_only():
    from maze_image_processing_pipeline_b200 import _lib
    _lib.build()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from maze_image_processing_pipeline_b200 import _lib, isotropic
    with pytest.raises(_lib.MazeLibraryError):
        isotropic.isotropic_erosion(np.ones((4, 4), bool), 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "maze_image_processing_pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libmaze_oracle" not in text, f


def test_config_mirrors_match_reference_schema_golden():
    import dataclasses
    import json
    from maze_image_processing_pipeline_b200.stage import SegmentationPostprocessingConfig, ThresholdSegmentationConfig
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "config_schema.json")))
    ref = doc["SegmentationPostprocessingConfig"]
    mine = {f.name: f.default for f in dataclasses.fields(SegmentationPostprocessingConfig)}
    assert set(mine) == set(ref)
    for k, v in ref.items():
        assert mine[k] == v["default"], k
    assert [f.name for f in dataclasses.fields(ThresholdSegmentationConfig)] == list(doc["ThresholdSegmentationConfig"])
    assert doc["short_forms"]["postprocess: true"] == dataclasses.asdict(SegmentationPostprocessingConfig())


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from maze_image_processing_pipeline_b200.stage import gather_object_tables, shard_bounds
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
n_img = 11
lo, hi = shard_bounds(n_img, rank, world)
# each image i contributes (i % 3) objects; local image index in column 57
rows = []
for i in range(lo, hi):
    for l in range(i % 3):
        r = np.zeros(64); r[0] = l + 1; r[1] = 10 * i + l; r[57] = i - lo
        rows.append(r)
local = np.array(rows).reshape(-1, 64)
full = gather_object_tables(local, lo)
if rank == 0:
    want = [(i, l + 1, 10 * i + l) for i in range(n_img) for l in range(i % 3)]
    got = [(int(r[57]), int(r[0]), int(r[1])) for r in full]
    assert got == want, (got, want)
    print("GATHER_OK", len(got))
else:
    assert full is None
dist.destroy_process_group()
'''


def test_sharded_tables_are_concatenated_in_image_order_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29591", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK 10" in outs[0]


def test_regions_find_regions_recalc_metadata_and_features():
    """Host-side tail of the stage (loki/pipeline.py:589-625) on a table produced by the oracle."""
    import oracle
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    from maze_image_processing_pipeline_b200.regions import find_regions, objects_of, recalc_metadata
    from maze_image_processing_pipeline_b200.stage import StageResult
    lab = np.zeros((40, 60), np.int32)
    lab[5:15, 10:30] = 1          # 10 x 20 block
    lab[20:23, 50:58] = 2         # small block near the right edge
    lab[30:35, 2:6] = 4           # label 3 absent (removed by a filter): its row has area 0
    img = np.full(lab.shape, 7, np.uint8)
    img[lab == 2] = 200
    img[6, 11] = 0
    table = oracle.regionprops_table(lab, img)
    g = BatchGeometry([40], [60])
    res = StageResult(g, g.pack_host([(lab > 0).view(np.uint8)]), g.pack_host([lab], dtype=np.int32),
                      np.array([0, 4], np.int32), table)
    regs = list(find_regions(res, 0, padding=75, image=img))
    assert [r.label for r in regs] == [1, 2, 4]
    # padded slice: start clipped at 0, stop unclipped (FindRegions / _enlarge_slice semantics)
    assert regs[0].bbox == (0, 0, 15 + 75, 30 + 75)
    assert regs[0].image.shape == (40, 60) and regs[0].image.sum() == 200
    m = recalc_metadata(regs[1], {"sample": "s"}, "{sample}_{object_sequence}")
    # the reference unpacks bbox as (y0, x0, x1, y1): width = max_row - min_col, height = max_col - min_row
    y0, x0, x1, y1 = regs[1].bbox
    assert (m["object_posx"], m["object_posy"]) == (x0, y0)
    assert m["object_width"] == x1 - x0 and m["object_height"] == y1 - y0
    assert m["object_width"] == (23 + 75) - 0 and m["object_id"] == "s_2"
    assert recalc_metadata(regs[0], {})["object_frac_invalid"] == 1 / 200
    # min_intensity drops regions whose brightest pixel is darker
    assert [r.label for r in find_regions(res, 0, min_intensity=100)] == [2]
    objs = objects_of(res, 0, meta={"frame": 3}, padding=0, image=img)
    assert len(objs) == 3 and objs[0]["frame"] == 3
    o = objs[0]
    assert o["object_area"] == 200 and o["object_bx"] == 10 and o["object_by"] == 5
    assert o["object_width"] == 20 and o["object_height"] == 10   # ZooProcess overrides the swapped values
    assert abs(o["object_x"] - 19.5) < 1e-12 and abs(o["object_y"] - 9.5) < 1e-12
    assert abs(o["object_angle"] - (table[0, oracle.F_ORIENT] / np.pi * 180 + 90)) < 1e-12
    assert o["object_intden"] == 200 * o["object_mean"] and o["object_range"] == 7.0


def test_regions_zooprocess_keys_with_shape_table():
    """The full ZooProcess key set from the moment table + the shape table (both produced by the oracle here)."""
    import oracle
    from oracle import shape as oshape
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    from maze_image_processing_pipeline_b200.regions import extract_roi, find_regions, objects_of
    from maze_image_processing_pipeline_b200.stage import StageResult
    lab = np.zeros((30, 40), np.int32)
    lab[4:14, 5:25] = 1
    lab[7:10, 9:15] = 0                 # a hole of 18 pixels
    lab[20:25, 30:33] = 2
    img = np.full(lab.shape, 50, np.uint8)
    table = oracle.regionprops_table(lab, img)
    g = BatchGeometry([30], [40])
    res = StageResult(g, g.pack_host([(lab > 0).view(np.uint8)]), g.pack_host([lab], dtype=np.int32),
                      np.array([0, 2], np.int32), table)
    res.shape_table = oshape.label_shape(lab)
    o1, o2 = objects_of(res, 0, padding=0, image=img)
    assert o1["object_area_exc"] == 200 - 18 and o1["object_area"] == 200 and o1["object_convex_area"] == 200
    assert abs(o1["object_%area"] - 18 / 200) < 1e-15 and o1["object_euler_number"] == 0
    assert o1["object_perim."] == oshape.perimeter(lab[4:14, 5:25] == 1)
    assert abs(o1["object_circ."] - 4 * np.pi * 200 / o1["object_perim."] ** 2) < 1e-12
    assert o1["object_solidity"] == 182 / 200 and o1["object_extent"] == 182 / 200
    assert o1["object_bounding_box_area"] == 200 and o1["object_intden"] == 200 * 50.0
    assert abs(o1["object_equivalent_diameter"] - np.sqrt(4 * 182 / np.pi)) < 1e-12
    assert o2["object_area"] == 15 and o2["object_euler_number"] == 1 and o2["object_solidity"] == 1.0
    assert abs(o2["object_local_centroid_row"] - 2.0) < 1e-12 and abs(o2["object_local_centroid_col"] - 1.0) < 1e-12
    # without the shape table the perimeter-based keys are absent and area falls back to the pixel count
    res.shape_table = None
    p1 = objects_of(res, 0, padding=0, image=img)[0]
    assert "object_perim." not in p1 and p1["object_area"] == 182
    # ExtractROI with apply_mask = False (the schema default) is the padded crop; alpha = 1 paints the rest
    reg = list(find_regions(res, 0, padding=3, image=img))[1]
    crop = extract_roi(img, reg)
    assert crop.shape == (5 + 6, 3 + 6) and crop.base is not None
    masked = extract_roi(img, reg, alpha=1, bg_color=7)
    assert (masked[reg.image] == 50).all() and (masked[~reg.image] == 7).all()
    # keep_background (the schema's default when apply_mask is on): only the pixels of OTHER objects are painted
    reg0 = list(find_regions(res, 0, padding=40, image=img))[1]      # a crop that reaches the first object too
    kept = extract_roi(img, reg0, alpha=1, bg_color=7, keep_background=True)
    lab = reg0.label_image
    assert ((lab != 0) & (lab != reg0.label)).any()
    assert np.array_equal(kept, np.where((lab == 0) | (lab == reg0.label), img[reg0.slice], 7))


def test_c_abi_rejects_bad_arguments_before_touching_the_gpu():
    """Entry points validate their arguments first (no CUDA call is made for a rejected request)."""
    from maze_image_processing_pipeline_b200 import _lib
    lib = _lib.lib()
    assert lib.maze_label_shape(None, None, None, None, 0, None, 0, 0, 0, 0, None, None, None) == _lib.MAZE_OK  # nothing to do
    assert lib.maze_label_shape(None, None, None, None, 5, None, 0, 0, 0, 0, None, None, None) == _lib.MAZE_ERR_BADARG
    with pytest.raises(ValueError):
        _lib.check(_lib.MAZE_ERR_BADARG, "maze_label_shape")
    assert lib.maze_count_scan(None, -1, None, None) == _lib.MAZE_ERR_BADARG
    # the graph variant of the step needs somewhere to put its handle
    import ctypes
    a = _lib.StepArgs()
    assert lib.maze_stage_step_graph(ctypes.byref(a), None, None, None, None) == _lib.MAZE_ERR_BADARG
    assert lib.maze_graph_destroy(None) == _lib.MAZE_OK
    # merge_labels: an explicit index needs its offsets, the cluster size is 1 or 8
    assert lib.maze_merge_labels_ex(None, None, None, 1, None, 1, ctypes.c_void_p(8), None, 1, 10.0, 5.0, None, None, None, None,
                                    None, None, None, None, None, 0, 1, None) == _lib.MAZE_ERR_BADARG
    assert lib.maze_merge_labels_ex(None, None, None, 1, None, 1, None, None, 1, 10.0, 5.0, None, None, None, None,
                                    None, None, None, None, None, 0, 4, None) == _lib.MAZE_ERR_BADARG


def test_crosses_footprints_and_minkowski_collapse():
    """disk(r, decomposition="crosses") as the live pipeline passes it (loki/pipeline.py:408-427): its expansion is
    the closed disk exactly for the radii SURVEY.md section 0.3 lists; a footprint sequence applied element by
    element (what skimage does) equals ONE pass with the collapsed footprint, borders included (scipy both)."""
    from scipy import ndimage as ndi
    from oracle import scipy_chain
    from maze_image_processing_pipeline_b200 import morphology as M
    diff = {}
    for r in range(1, 21):
        F, D = M.collapse(M.disk(r, decomposition="crosses")), M.disk(r).astype(bool)
        assert F.shape == D.shape
        diff[r] = int((F ^ D).sum())
    assert [r for r in diff if diff[r] == 0] == [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 16, 19, 20]
    assert (diff[13], diff[17], diff[18]) == (8, 16, 8)
    assert M.chord_table(M.disk(2, decomposition="crosses")).tolist() == [2, 1, 0]
    assert M.chord_table(np.ones((3, 7), np.uint8)).tolist() == [3, 3]
    for bad in (np.array([[1, 0, 1], [0, 1, 0], [1, 0, 1]], np.uint8), np.ones((2, 3), np.uint8),
                np.array([[0, 1, 0], [0, 1, 0], [1, 1, 1]], np.uint8)):
        with pytest.raises(NotImplementedError):
            M.chord_table(bad)
    rng = np.random.default_rng(4)
    for r in (1, 2, 3, 5, 13):
        seq = M.disk(r, decomposition="crosses")
        fp = M.collapse(seq)
        for shape, p in (((40, 57), 0.5), ((9, 64), 0.9), ((33, 5), 0.97)):
            m = rng.random(shape) < p
            assert np.array_equal(scipy_chain.binary_erosion(m, seq), ndi.binary_erosion(m, structure=fp, border_value=1))
            assert np.array_equal(scipy_chain.binary_dilation(m, seq), ndi.binary_dilation(m, structure=fp))


def test_ecotaxa_writer_and_rescale_max_intensity(tmp_path):
    """Archive layout of the output stage (loki/pipeline.py:1225-1236) and rescale_max_intensity (:382-383)."""
    import zipfile
    from PIL import Image
    from maze_image_processing_pipeline_b200.ecotaxa import EcotaxaWriter, rescale_max_intensity
    rng = np.random.default_rng(0)
    img = rng.integers(0, 120, size=(20, 30), dtype=np.uint8)
    out = rescale_max_intensity(img)
    assert out.dtype == np.uint8 and out.max() == 255 and out.min() == 0 * img.min()
    assert np.array_equal(out, np.asarray(img / img.max() * 255, dtype=np.uint8))
    assert np.array_equal(rescale_max_intensity(np.zeros((3, 3), np.uint8)), np.zeros((3, 3), np.uint8))
    fn = str(tmp_path / "export.zip")
    mask = img > 60
    with EcotaxaWriter(fn, store_types=True) as w:
        w.add([("obj_1.png", img), ("obj_1_mask.png", mask)], {"object_id": "obj_1", "object_area": 12.0, "object_label": 1})
        w.add([("obj_2.png", out)], {"object_id": "obj_2", "object_area": 7.5, "object_label": 2, "object_note": "x"})
    with zipfile.ZipFile(fn) as z:
        assert sorted(z.namelist()) == ["ecotaxa_export.tsv", "obj_1.png", "obj_1_mask.png", "obj_2.png"]
        rows = z.read("ecotaxa_export.tsv").decode().rstrip("\n").split("\n")
        assert rows[0].split("\t") == ["img_file_name", "img_rank", "object_id", "object_area", "object_label", "object_note"]
        assert rows[1].split("\t") == ["[t]", "[f]", "[t]", "[f]", "[f]", "[t]"]
        assert rows[2].split("\t") == ["obj_1.png", "1", "obj_1", "12.0", "1", ""]
        assert rows[3].split("\t")[:2] == ["obj_1_mask.png", "2"] and rows[4].split("\t")[-1] == "x"
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(z.read("obj_1.png")))), img)
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(z.read("obj_1_mask.png")))) > 0, mask)
    with pytest.raises(ValueError):
        with EcotaxaWriter(str(tmp_path / "bad.zip")) as w:
            w.add([("obj.xyz", img)], {})


def test_zooprocess_keys_follow_the_padded_region_like_the_reference():
    """The reference hands the SAME padded region (FindRegions(padding=75)) to recalc_metadata and to
    CalculateZooProcessFeatures (loki/pipeline.py:589-625): bbox-derived keys come from the padded slice (start clipped
    at 0, stop not), bbox_area / extent / local_centroid from region.image, i.e. the slice clipped to the frame."""
    import oracle
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    from maze_image_processing_pipeline_b200.regions import objects_of
    from maze_image_processing_pipeline_b200.stage import StageResult
    lab = np.zeros((200, 260), np.int32)
    lab[100:120, 90:130] = 1       # padding fits on every side
    lab[10:30, 240:255] = 2        # padding is cut at the top and at the right edge
    lab[150:160, 5:25] = 5         # labels 3, 4 absent: sequence numbers close the gap
    img = np.full(lab.shape, 9, np.uint8)
    table = oracle.regionprops_table(lab, img)
    g = BatchGeometry([200], [260])
    res = StageResult(g, g.pack_host([(lab > 0).view(np.uint8)]), g.pack_host([lab], dtype=np.int32),
                      np.array([0, 5], np.int32), table)
    pad = 75
    objs = objects_of(res, 0, padding=pad, image=img)
    assert [o["object_sequence"] for o in objs] == [1, 2, 3] and [o["object_label"] for o in objs] == [1, 2, 3]
    for o, l in zip(objs, (1, 2, 5)):
        rows, cols = np.nonzero(lab == l)
        sl = (slice(max(0, rows.min() - pad), rows.max() + 1 + pad), slice(max(0, cols.min() - pad), cols.max() + 1 + pad))
        image = lab[sl] == l                                   # RegionProperties.image: numpy clips the slice
        bbox = (sl[0].start, sl[1].start, sl[0].stop, sl[1].stop)  # RegionProperties.bbox: the slice as given
        assert o["object_bx"] == bbox[1] and o["object_by"] == bbox[0]
        assert o["object_width"] == bbox[3] - bbox[1] and o["object_height"] == bbox[2] - bbox[0]
        assert o["object_bounding_box_area"] == image.size
        assert o["object_extent"] == image.sum() / image.size
        assert abs(o["object_local_centroid_row"] - (rows.mean() - sl[0].start)) < 1e-9
        assert abs(o["object_local_centroid_col"] - (cols.mean() - sl[1].start)) < 1e-9
        assert abs(o["object_y"] - rows.mean()) < 1e-9 and abs(o["object_x"] - cols.mean()) < 1e-9
        assert o["object_posx"] == bbox[1] and o["object_posy"] == bbox[0]


def _runs_of(lab, rpb):
    """Run list + band table of a label image the way maze_band_stage lays them out (test helper)."""
    from maze_image_processing_pipeline_b200.device import BAND_OUT_DTYPE, RUN_DTYPE
    h, w = lab.shape
    runs, band_out = [], []
    for y0 in range(0, h, rpb):
        base = len(runs)
        for y in range(y0, min(h, y0 + rpb)):
            x = 0
            while x < w:
                if lab[y, x]:
                    x1 = x
                    while x1 + 1 < w and lab[y, x1 + 1] == lab[y, x]:
                        x1 += 1
                    runs.append((y, x, x1, lab[y, x]))
                    x = x1 + 1
                else:
                    x += 1
        band_out.append((base, len(runs) - base, 0, 0))
    return np.array(runs, RUN_DTYPE).reshape(-1), np.array(band_out, BAND_OUT_DTYPE).reshape(-1)


def test_host_expand_and_crops_from_run_lists():
    """maze_host_expand / maze_host_expand_crop (the host half of the compact result transport) and the lazy
    StageResult built on them, on hand-made run lists -- no GPU involved."""
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    from maze_image_processing_pipeline_b200.regions import find_regions
    from maze_image_processing_pipeline_b200.stage import StageResult
    rng = np.random.default_rng(5)
    labs, rpbs = [], [7, 64, 3, 1]
    for (h, w) in [(40, 70), (33, 1), (1, 90), (25, 31)]:
        lab = (rng.random((h, w)) < 0.4) * rng.integers(1, 6, (h, w))
        labs.append(lab.astype(np.int32))
    g = BatchGeometry([l.shape[0] for l in labs], [l.shape[1] for l in labs])
    all_runs, all_bo, band_off = [], [], [0]
    for lab, rpb in zip(labs, rpbs):
        r, bo = _runs_of(lab, rpb)
        bo["base"] += sum(len(x) for x in all_runs)
        all_runs.append(r)
        all_bo.append(bo)
        band_off.append(band_off[-1] + len(bo))
    runs = np.concatenate(all_runs)
    res = StageResult.from_runs(g, runs, np.concatenate(all_bo), np.asarray(band_off, np.int32),
                                np.asarray(rpbs, np.int32), {}, np.zeros(5, np.int32), np.zeros((0, 64)))
    assert res.compact
    for i, lab in enumerate(labs):
        assert np.array_equal(res.labels(i), lab) and np.array_equal(res.mask(i), lab > 0)
        assert res.labels(i).dtype == np.int32 and res.mask(i).dtype == bool
        rr = res.runs(i)
        assert int((rr["x1"].astype(int) - rr["x0"] + 1).sum()) == int((lab > 0).sum())
        h, w = lab.shape
        for sl in [(slice(0, h), slice(0, w)), (slice(h // 3, h), slice(w // 4, w // 2 + 1)), (slice(2, 2), slice(0, w)),
                   (slice(0, h + 75), slice(max(0, w - 5), w + 75))]:  # stops beyond the frame are clipped like numpy
            assert np.array_equal(res.object_mask(i, sl), (lab > 0)[sl])
            for l in (1, 3, 9):
                assert np.array_equal(res.object_mask(i, sl, l), lab[sl] == l)
    # regions of a compact result expand one object at a time
    import oracle
    res.table = oracle.regionprops_table(labs[0], np.full(labs[0].shape, 5, np.uint8))
    res.lab_off = np.array([0, len(res.table)] + [len(res.table)] * 3, np.int32)
    for reg in find_regions(res, 0, padding=4):
        assert np.array_equal(reg.image, labs[0][reg.slice] == reg.label)
    dense = res.materialize(threads=3)
    assert not dense.compact
    for i, lab in enumerate(labs):
        assert np.array_equal(dense.labels(i), lab) and np.array_equal(dense.mask(i), lab > 0)


def test_host_expand_streams_large_vignettes_and_survives_unordered_runs():
    """The expansion composes the output chunk by chunk (runs in raster order, chunks of 8192 pixels leave with
    non-temporal stores): vignettes of several chunks with runs that cross chunk borders, the same run list in
    shuffled order (falls back to painting in place) and a run outside the vignette (dropped, never written)."""
    from maze_image_processing_pipeline_b200.device import BatchGeometry, RUN_DTYPE
    from maze_image_processing_pipeline_b200.stage import StageResult
    rng = np.random.default_rng(11)
    labs = []
    for (h, w) in [(300, 257), (64, 1000), (513, 64)]:
        lab = np.zeros((h, w), np.int32)
        for _ in range(40):  # long runs: many cross a chunk border
            y0, x0 = int(rng.integers(0, h - 3)), int(rng.integers(0, w - 3))
            lab[y0:y0 + int(rng.integers(1, 30)), x0:x0 + int(rng.integers(1, w))] = int(rng.integers(1, 9))
        labs.append(lab)
    g = BatchGeometry([l.shape[0] for l in labs], [l.shape[1] for l in labs])
    for shuffle in (False, True):
        all_runs, all_bo, band_off, rpbs = [], [], [0], [50, 64, 100]
        for lab, rpb in zip(labs, rpbs):
            r, bo = _runs_of(lab, rpb)
            if shuffle:  # inside every band
                for b in bo:
                    seg = r[b["base"]:b["base"] + b["n_runs"]]
                    seg[:] = seg[rng.permutation(len(seg))]
            bo["base"] += sum(len(x) for x in all_runs)
            all_runs.append(r)
            all_bo.append(bo)
            band_off.append(band_off[-1] + len(bo))
        runs = np.concatenate(all_runs)
        if shuffle:  # and one run that lies outside its vignette: it must not be written anywhere
            bad = np.zeros(1, RUN_DTYPE)
            bad["y"], bad["x0"], bad["x1"], bad["label"] = 60000, 0, 5, 3
            k = int(all_bo[0][0]["base"])
            runs[k] = bad[0]
        res = StageResult.from_runs(g, runs, np.concatenate(all_bo), np.asarray(band_off, np.int32),
                                    np.asarray(rpbs, np.int32), {}, np.zeros(4, np.int32), np.zeros((0, 64)))
        dense = res.materialize(threads=2)
        for i, lab in enumerate(labs):
            got = dense.labels(i)
            if shuffle and i == 0:  # (the overwritten run of vignette 0 is missing, nothing else differs)
                assert ((got == lab) | (got == 0)).all() and (got != lab).sum() <= labs[0].shape[1]
            else:
                assert np.array_equal(got, lab), (shuffle, i)
                assert np.array_equal(dense.mask(i), lab > 0)


def test_band_plan_covers_every_row_once_and_fits_the_planes():
    """Band plan of maze_band_stage: the bands of a vignette are consecutive, cover its rows exactly once with a uniform
    height, and a band plus its halo never exceeds MAZE_BAND_PLANE_WORDS; vignettes that cannot be cut are listed."""
    from maze_image_processing_pipeline_b200._lib import BAND_PLANE_WORDS
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    rng = np.random.default_rng(0)
    hs = np.concatenate([rng.integers(1, 1500, 200), [1, 70000, 4096, 300, 2, 5000]])
    ws = np.concatenate([rng.integers(1, 1500, 200), [70000, 1, 4096, 20000, 65535, 40]])
    g = BatchGeometry(hs, ws)
    for halo in (0, 4, 6, 40):
        bands, off, left = g.band_plan(halo)
        assert off[0] == 0 and off[-1] == len(bands)
        for i in range(g.n_img):
            b = bands[off[i]:off[i + 1]]
            wpr = (int(ws[i]) + 31) // 32
            if len(b) == 0:
                assert i in left
                assert hs[i] >= 65536 or ws[i] >= 65536 or BAND_PLANE_WORDS // wpr - 2 * halo < max(1, halo)
                continue
            assert i not in left and (b["img"] == i).all()
            assert b["y0"][0] == 0 and b["y1"][-1] == hs[i] and (b["y0"][1:] == b["y1"][:-1]).all()
            assert (b["rpb"] == b["rpb"][0]).all() and (b["y0"] == np.arange(len(b)) * b["rpb"][0]).all()
            if len(b) == 1:
                assert hs[i] * wpr <= BAND_PLANE_WORDS                      # the whole vignette, no halo
            else:
                lo = np.maximum(0, b["y0"] - halo)
                hi = np.minimum(int(hs[i]), b["y1"] + halo)
                assert ((hi - lo) * wpr <= BAND_PLANE_WORDS).all()


def test_band_plan_cuts_a_lone_frame_into_enough_bands_for_every_sm():
    """A batch that is one frame gets thinner bands (down to twice the halo) so that ~3 CTAs per SM exist; the
    invariants of the plan stay; a batch with plenty of bands is left alone."""
    from maze_image_processing_pipeline_b200._lib import BAND_PLANE_WORDS
    from maze_image_processing_pipeline_b200.device import BAND_TARGET_CTAS, BatchGeometry
    for halo in (0, 6):
        g = BatchGeometry([4096], [4096])
        bands, off, left = g.band_plan(halo)
        assert len(left) == 0 and 300 <= len(bands) <= BAND_TARGET_CTAS + 64
        assert bands["y0"][0] == 0 and bands["y1"][-1] == 4096 and (bands["y0"][1:] == bands["y1"][:-1]).all()
        assert (bands["rpb"] == bands["rpb"][0]).all() and bands["rpb"][0] >= max(2 * halo, 4)
        assert ((np.minimum(4096, bands["y1"] + halo) - np.maximum(0, bands["y0"] - halo)) * 128 <= BAND_PLANE_WORDS).all()
    many = BatchGeometry([4096] * 8, [4096] * 8)       # enough bands already: full-height bands
    bands, off, _ = many.band_plan(6)
    assert bands["rpb"][0] == BatchGeometry([4096] * 8, [4096] * 8).band_plan(6)[0]["rpb"][0] > 24


def test_host_pack_async_equals_sync():
    """maze_host_pack_start / maze_host_pack_wait (background packing of the next batch) against maze_host_pack."""
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (int(h), int(w)), dtype=np.uint8) for h, w in rng.integers(1, 700, (300, 2))]
    imgs.append(np.asfortranarray(rng.integers(0, 256, (50, 60), dtype=np.uint8)))  # not C-contiguous: copied first
    g = BatchGeometry.from_images(imgs)
    want = g.pack_host(imgs, out=np.zeros(g.total_px, np.uint8), threads=3)
    got = np.zeros(g.total_px, np.uint8)
    wait = g.pack_host_start(imgs, out=got, threads=5)
    wait()
    wait()  # idempotent
    assert np.array_equal(got, want)
    for i, im in enumerate(imgs):
        assert np.array_equal(g.view(got, i), im)


def test_host_array_pool_never_hands_out_an_array_that_is_still_referenced():
    """materialize() recycles its flat host arrays: only arrays nobody refers to any more (no view, no result)."""
    from maze_image_processing_pipeline_b200.stage import _HostArrayPool
    pool = _HostArrayPool(keep=2)
    a = pool.take(1000, np.uint8)
    base_a = a.base
    view = a[100:200]          # what StageResult.mask(i) hands to the caller
    del a
    b = pool.take(1000, np.uint8)
    assert b.base is not base_a            # the view keeps the first array busy
    view[:] = 7
    b[:] = 1
    assert (view == 7).all()
    del view, b
    c = pool.take(900, np.uint8)           # both are idle now: one of them comes back
    assert c.base is base_a or c.size == 900
    assert len(pool._bufs) <= 2
    d = pool.take(10, np.int32)            # another dtype never aliases
    assert d.dtype == np.int32 and d.base is not c.base
    for _ in range(5):                     # busy arrays beyond `keep` are simply not pooled
        pool.take(50, np.uint8)
    assert len(pool._bufs) <= 2


def test_zooprocess_table_equals_the_per_object_dicts():
    """The batch-wide, vectorised metadata table against objects_of (FindRegions -> recalc_metadata ->
    CalculateZooProcessFeatures per object), with and without the shape table, padding and min_intensity."""
    import oracle
    from oracle import shape as oshape
    from maze_image_processing_pipeline_b200.device import BatchGeometry
    from maze_image_processing_pipeline_b200.regions import objects_of, zooprocess_table
    from maze_image_processing_pipeline_b200.stage import StageResult
    rng = np.random.default_rng(5)
    labs, imgs = [], []
    for k in range(5):
        h, w = int(rng.integers(30, 90)), int(rng.integers(30, 90))
        m = rng.random((h, w)) < 0.08
        m[h // 4:h // 2, w // 5:w // 2] = True
        lab = oracle.label(m)[0]
        if k == 2:
            lab[lab == 2] = 0                    # a gap in the numbering, as the label filters leave it
        labs.append(lab)
        imgs.append(rng.integers(0, 200, (h, w)).astype(np.uint8))
    g = BatchGeometry.from_images(labs)
    tables = [oracle.regionprops_table(l, im) for l, im in zip(labs, imgs)]
    off = np.concatenate([[0], np.cumsum([len(t) for t in tables])]).astype(np.int32)
    res = StageResult(g, g.pack_host([(l > 0).view(np.uint8) for l in labs]), g.pack_host(labs, dtype=np.int32), off,
                      np.concatenate(tables))
    for with_shape in (False, True):
        res.shape_table = np.concatenate([oshape.label_shape(l) for l in labs]) if with_shape else None
        for padding, min_int in ((75, None), (0, None), (3, 150)):
            cols = zooprocess_table(res, padding=padding, min_intensity=min_int)
            want = [(i, o) for i in range(len(labs)) for o in objects_of(res, i, padding=padding, min_intensity=min_int,
                                                                         image=imgs[i])]
            assert len(cols["image_index"]) == len(want) > 10
            for j, (i, o) in enumerate(want):
                assert cols["image_index"][j] == i
                for key, v in o.items():
                    got = cols[key][j]
                    if isinstance(v, float) and v != v:
                        assert got != got, key
                    else:
                        assert got == v, (key, got, v)
                assert set(o) <= set(cols)
