"""CPU tests: host-side mirror of the reference rules vs golden vectors from the imported reference,
the C-ABI export list vs include/synseg.h, data model, synthetic generator, and the world_size-2 gloo path."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from synapta_image_segmentation_b200 import geometry as G  # noqa: E402
from synapta_image_segmentation_b200.datamodel import (BoundingBox, ChartSpecificData, FigureSpecificData, VisualSegment,  # noqa: E402
                                                       VisualType)


def BB(*a):
    return BoundingBox(*a, 612, 792)


@pytest.fixture(scope="module")
def geo():
    return json.load(open(os.path.join(GOLD, "reference_geometry.json")))


def test_overlap_rules(geo):
    assert G.calculate_overlap_ratio(BB(0, 0, 10, 10), BB(5, 5, 20, 20)) == geo["overlap_ratio"] == 0.25
    ex = [{"bbox": BB(0, 0, 100, 100)}]
    got = [G.overlaps_with_existing(BB(50, 0, 150, 100), ex), G.overlaps_with_existing(BB(49, 0, 149, 100), ex)]
    assert got == geo["overlaps_existing"] == [False, True]
    assert G.find_conflicting(BB(0, 0, 10, 10), [type("S", (), {"bbox": BB(5, 5, 20, 20)})()]) is None      # 0.25 <= 0.4
    assert G.find_conflicting(BB(0, 0, 10, 10), [type("S", (), {"bbox": BB(2, 2, 20, 20)})()]) is not None  # 0.64


def test_drawing_distance_and_clusters(geo):
    assert [G.drawing_distance([0, 0, 10, 10], [10, 10, 20, 20]), G.drawing_distance([0, 0, 10, 10], [13, 14, 20, 20])] == geo["drawing_distance"]
    rects5 = [[x, 0, x + 10, 10] for x in (0, 60, 150, 240, 330)]
    assert [[rects5[i] for i in c] for c in G.cluster_rects(rects5)] == geo["cluster_5"] == []
    rects = geo["cluster_random_in"]
    assert [[rects[i] for i in c] for c in G.cluster_rects(rects)] == geo["cluster_random_out"]


def test_regions_from_rects(geo):
    regs = G.regions_from_rects([[100 + 5 * i, 100, 110 + 5 * i, 300] for i in range(6)], 612.0, 792.0)
    got = [dict(bbox=[r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1], caption=r["caption"],
                detection_method=r["detection_method"], notes=r["notes"]) for r in regs]
    assert got == geo["detect_by_drawings"]
    regs = G.regions_from_rects(geo["cluster_random_in"], 612.0, 792.0)
    got = [dict(bbox=[r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1], notes=r["notes"]) for r in regs]
    assert got == geo["detect_by_drawings_random"]


def test_validate_region_bit_identical_scores(geo):
    imgs_ = {"noise": (400, 300, float(np.var(_gray("noise")))), "flat": (300, 300, 0.0), "small": (300, 40, float(np.var(_gray("small"))))}
    for v in geo["validate"]:
        w, h, var = imgs_[v["image"]]
        score, notes = G.validate_region(BB(*v["bbox"]), w, h, var, 792.0)
        assert score == v["score"] and notes == v["notes"], v      # == on floats: same f64 operation order


def _gray(name):
    from PIL import Image
    if name == "noise":
        a = np.random.default_rng(1).integers(0, 256, (300, 400, 3), dtype=np.uint8)
    else:
        a = np.random.default_rng(2).integers(0, 256, (40, 300, 3), dtype=np.uint8)
    return np.array(Image.fromarray(a).convert("L"))


def test_resolve_conflict(geo):
    var = {"noise": float(np.var(_gray("noise"))), "flat": 0.0}
    for v in geo["resolve_conflict"]:
        d, why = G.resolve_conflict(BB(*v["emb"]), v["confidence"], var[v["image"]], BB(*v["cap"]), v["caption"])
        assert (d, why) == (v["decision"], v["reasons"])


def test_pass2_conflict_resolution_matches_reference(geo):
    """`_extract_images_from_page` pass 2 (pdf_image_segmentation.py:2822-2847): which segments survive when validated candidates
    meet caption-based ones -- decisions recorded from the reference's own _find_conflicting_segment / _resolve_conflict."""
    var = {"noise": float(np.var(_gray("noise"))), "flat": 0.0}
    for sc in geo["pass2"]:
        caps = [dict(bbox=BB(*c["bbox"]), caption=c["caption"], detection_method="caption_based", confidence=0.9) for c in sc["captions"]]
        cands = [dict(bbox=BB(*c["bbox"]), confidence=c["confidence"], variance=var[c["image"]], detection_method="raster_cc") for c in sc["candidates"]]
        out = G.resolve_page_conflicts(caps, cands, sc["drawings"])
        got = [dict(method="caption_based" if r["detection_method"] == "caption_based" else "embedded_image",
                    bbox=[r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1]) for r in out]
        assert got == sc["result"], sc


def test_vectorised_clustering_equals_the_reference_loops():
    """cluster_rects (numpy inner loop) == the reference's two nested loops, also on grids full of pairs at exactly 100 pt."""
    rng = np.random.default_rng(3)
    for trial in range(60):
        n = int(rng.integers(16, 220))
        if trial % 3 == 0:
            xy = rng.uniform(0, 700, (n, 2)); wh = rng.uniform(1, 60, (n, 2))
        elif trial % 3 == 1:
            xy = 0.24 * rng.integers(0, 3000, (n, 2)); wh = 0.24 * rng.integers(1, 200, (n, 2))
        else:
            xy = 20.0 * rng.integers(0, 40, (n, 2)); wh = 20.0 * rng.integers(1, 3, (n, 2))      # (60, 80) gaps: distance exactly 100
        rects = np.concatenate([xy, xy + wh], 1).tolist()
        assert G.cluster_rects(rects) == G._cluster_rects_scalar(rects, G.CLUSTER_GAP, G.CLUSTER_MIN_MEMBERS)


def test_merge_visual_regions():
    a = {"bbox": BB(0, 0, 100, 100), "caption_bbox": [10, 120, 90, 130]}
    dup = {"bbox": BB(10, 10, 90, 90)}                  # 100 % of its area inside a -> dropped
    above_caption = {"bbox": BB(0, 300, 100, 400)}
    near_caption = {"bbox": BB(5, 20, 95, 115)}         # caption top-left within +-50 pt of its bottom edge... and duplicate
    far = {"bbox": BB(300, 300, 400, 400)}
    out = G.merge_visual_regions([a], [dup, above_caption, near_caption, far])
    assert out == [a, above_caption, far]


def test_datamodel_to_dict_schema():
    seg = VisualSegment("textbook_001_p000_61f12f4c", VisualType.IMAGE, "textbook_001", 1, BoundingBox(106.37, 541.88, 439.38, 749.38, 651.6, 795.6),
                        image_bytes=b"x", confidence=np.float64(1.0),
                        notes="Validation: good_size, substantial_dimensions, good_aspect_ratio, good_position, good_content_variance")
    seg.chart_data = ChartSpecificData(grid_detected=np.bool_(True), estimated_data_points=np.int64(23), color_scheme=["#234665"])
    seg.figure_data = FigureSpecificData(contains_chart=True)
    d = seg.to_dict()
    ref = json.load(open("/root/reference/extracted_visuals_excelSS/textbook_001_visual_segments.json"))["segments"][0] \
        if os.path.exists("/root/reference/extracted_visuals_excelSS/textbook_001_visual_segments.json") else None
    if ref is not None:     # same key set as the reference's shipped JSON record (build container only)
        assert set(ref) - set(d) <= {"image_details"} and set(ref["bbox"]) == set(d["bbox"])
    assert "image_bytes" not in d and d["segment_type"] == "image" and d["bbox"]["width"] == pytest.approx(333.01)
    assert d["chart_details"]["has_grid"] is True and d["chart_details"]["data_points"] == 23
    assert d["figure_details"]["contains_chart"] is True
    json.dumps(d)
    assert BoundingBox(72, 72, 144, 216, 612, 792).to_pixels(300) == (300, 300, 300, 600)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports exactly the functions include/synseg.h declares."""
    from synapta_image_segmentation_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "synseg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(synseg_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (synseg_\w+)", out))
    assert exported == declared, exported ^ declared
    assert lib.synseg_version() == 200
    # the ctypes signatures carry as many arguments as the prototypes in the header
    for m in re.finditer(r"\b(synseg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n_params = 0 if params in ("", "void") else params.count(",") + 1
        assert n_params == len(_lib.SIGNATURES[name][1]), (name, n_params, len(_lib.SIGNATURES[name][1]))
    import torch
    if not torch.cuda.is_available():      # fails loudly without a GPU: no CPU fallback
        import ctypes as C
        h = C.c_void_p()
        assert lib.synseg_create(0, C.byref(h)) != 0 and lib.synseg_last_error()


def test_synth_is_deterministic_and_shaped():
    from synapta_image_segmentation_b200.synth import page_shape, synth_page
    assert page_shape(300) == (3300, 2550) and page_shape(150) == (1650, 1275)
    a, ta = synth_page(5, 72, n_figures=2)
    b, tb = synth_page(5, 72, n_figures=2)
    assert np.array_equal(a, b) and ta == tb and a.shape == (792, 612, 3) and len(ta) == 2


def test_shard_and_keys():
    from synapta_image_segmentation_b200.dedup import region_key, shard_pages
    for n, w in [(8000, 8), (1000, 3), (5, 8)]:
        pages = [p for r in range(w) for p in shard_pages(n, r, w)]
        assert pages == list(range(n))
    assert region_key(3, 2) == (3 << 16) | 2
    with pytest.raises(ValueError):
        region_key(1, 70000)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from synapta_image_segmentation_b200.dedup import gather_hashes, shard_pages, region_key
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank hashes "its" pages with a deterministic stand-in hash; page p and p+4 are duplicates
pages = list(shard_pages(10, rank, world))
h = torch.tensor([(p % 4) * 0x0101010101 + 7 for p in pages], dtype=torch.int64)
k = torch.tensor([region_key(p, 0) for p in pages], dtype=torch.int64)
hh, kk = gather_hashes(h, k, capacity=8)
assert kk.tolist() == [region_key(p, 0) for p in range(10)], kk.tolist()
assert hh.tolist() == [(p % 4) * 0x0101010101 + 7 for p in range(10)]
# replicated decision (numpy stand-in for the CUDA Hamming kernel): identical on both ranks
keep = [not any(kk[j] < kk[i] and bin(int(hh[i]) ^ int(hh[j])).count("1") <= 4 for j in range(len(kk))) for i in range(len(kk))]
assert keep == [True] * 4 + [False] * 6
out = [None] * world
dist.all_gather_object(out, keep)
assert all(o == keep for o in out)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_gather_and_replicated_dedup(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script), ROOT], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_writers_match_reference_bytes(tmp_path):
    """SegmentWriter == the reference's _initialize_json_file / _save_results / _save_summary_csv (S:3852-3952), byte
    for byte, on the constructed segments of tests/golden/make_schema_golden.py; key layout == the shipped run artefact."""
    import importlib.util
    from synapta_image_segmentation_b200 import datamodel as dm
    from synapta_image_segmentation_b200.writers import SegmentWriter, load_segments_json
    gold = json.load(open(os.path.join(GOLD, "reference_writers.json")))
    spec = importlib.util.spec_from_file_location("make_schema_golden_segments", os.path.join(GOLD, "make_schema_golden.py"))
    src = open(spec.origin).read()
    ns = {}
    exec(src[src.index("def segments(mod):"):src.index("def main():")], ns)      # only the segment constructor, no reference import
    segs = ns["segments"](dm)
    w = SegmentWriter("textbook_001", "book.pdf", tmp_path, flush_every=2)
    w.initialize()
    assert w.output_json.read_text(encoding="utf-8") == gold["initialized_json"]
    assert w.extend(segs) == len(segs)
    assert w.append(segs[0]) is False                       # duplicate id is written once (S:3886-3887)
    w.save_results()
    assert w.output_json.read_text(encoding="utf-8") == gold["json"]
    assert w.output_csv.read_bytes().decode("utf-8") == gold["csv"]
    doc = load_segments_json(w.output_json)
    assert list(doc.keys()) == gold["a1_top_keys"] and doc["total_segments"] == len(segs)
    assert list(doc["segments"][0]["bbox"].keys()) == gold["a1_bbox_keys"]
    image_seg = [s for s in doc["segments"] if s["segment_type"] == "image"][0]
    assert list(image_seg.keys()) == gold["a1_segment_keys"]                     # same keys, same order as the shipped run
    assert list(image_seg["image_details"].keys()) == gold["a1_image_details_keys"]
    assert list(image_seg["image_data"].keys()) == gold["a1_image_data_keys"]
    assert w.summary_csv_text().splitlines()[0] == gold["a1_csv_header"]


def test_candidate_regions_with_pdf_priors():
    """Host stage of the detector (no GPU): component boxes -> region dicts, merged with caption-based priors exactly
    like _detect_visual_regions merges drawing regions into caption regions (pdf_image_segmentation.py:3114-3144)."""
    from synapta_image_segmentation_b200.detector import DetectConfig, RasterRegionDetector
    det = RasterRegionDetector(DetectConfig(dpi=72), ctx=object())          # the host stage never touches the context
    # stats rows: x, y, w, h, area (px = points at 72 DPI); row 0 = background
    stats = np.array([[0, 0, 612, 792, 400000],
                      [100, 100, 200, 150, 30000],       # a figure-sized component
                      [100, 400, 220, 160, 35200],       # another one, 60 pt above a caption
                      [400, 600, 8, 8, 64], [415, 600, 8, 8, 64], [430, 600, 8, 8, 64]], np.int32)   # three specks: one cluster, too small to keep
    plain = det.candidate_regions(stats, len(stats), 612.0, 792.0)
    assert [(r["detection_method"], r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1) for r in plain] == \
        [("raster_cc", 100.0, 100.0, 300.0, 250.0), ("raster_cc", 100.0, 400.0, 320.0, 560.0)]
    cap1 = {"bbox": BB(90, 90, 310, 260), "caption": "Figure 1.1 Something", "caption_bbox": (100, 265, 300, 278),
            "detection_method": "caption_based", "notes": "Caption: Figure 1.1"}
    cap2 = {"bbox": BB(400, 100, 560, 220), "caption": "Exhibit 2", "caption_bbox": (120, 570, 300, 583),
            "detection_method": "caption_based", "notes": "Caption: Exhibit 2"}
    merged = det.candidate_regions(stats, len(stats), 612.0, 792.0, priors=[cap1, cap2])
    # both priors stay; component 1 duplicates cap1 (all of it lies inside); component 2 sits right above cap2's caption
    assert [r["detection_method"] for r in merged] == ["caption_based", "caption_based"]
    far = dict(cap2, caption_bbox=(120, 700, 300, 713))
    merged = det.candidate_regions(stats, len(stats), 612.0, 792.0, priors=[cap1, far])
    assert [r["detection_method"] for r in merged] == ["caption_based", "caption_based", "raster_cc"]
    assert (merged[2]["bbox"].x0, merged[2]["bbox"].y0) == (100.0, 400.0)
    with pytest.raises(RuntimeError):
        det.candidate_regions(stats, -7, 612.0, 792.0)                      # label overflow is an error on the host rules (the detector retries)
    # the live flow of the reference (pass 2 of _extract_images_from_page): a validated raster candidate against the caption-based segments
    cands = [dict(r, confidence=0.8, variance=50.0) for r in plain]
    out = G.resolve_page_conflicts([dict(cap1, confidence=0.9)], cands)
    # candidate 1 overlaps cap1 (ratio 1.0 > 0.4): caption 3 + 2 (larger) vs embedded 1 (confidence) -> caption stays; candidate 2 is added
    assert [r["detection_method"] for r in out] == ["caption_based", "raster_cc"] and out[1]["bbox"].y0 == 400.0


def test_decode_colors_row_layout():
    """Host decoding of one synseg_colors_crops row: { mask_px, k, (cluster pixels << 24 | R << 16 | G << 8 | B) x k, 0 ... }."""
    from synapta_image_segmentation_b200.hints import FeatureHints
    row = np.array([14900, 3, (6400 << 24) | 0x1EA03C, (4500 << 24) | 0x283CD2, (4000 << 24) | 0xC81E28, 0, 0], np.int64)
    assert FeatureHints.decode_colors(row) == dict(mask_px=14900, dominant_colors=["#1ea03c", "#283cd2", "#c81e28"],
                                                   color_weights=[6400, 4500, 4000])
    assert FeatureHints.decode_colors(np.array([99, 0, 0, 0, 0, 0, 0], np.int64)) == dict(mask_px=99, dominant_colors=[], color_weights=[])


def test_committed_ncu_traffic_table_feeds_the_bench_roofline():
    """bench.py takes roofline.traffic / binding_resource / issue_slots of the dominant kernel from profiles/traffic.json: the
    table must name the kernels of the page pipeline with the fields bench.py reads, for the 50-page launches the bench profiles."""
    import json
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert tj["pages_per_launch"] == 50
    for k in ("canny_rgb", "adaptive_mean", "bitmorph_h", "bitmorph_v", "ccl_merge", "ccl_final"):
        e = tj["kernels"][k]
        assert e["dram_read_bytes"] > 0 and e["dram_write_bytes"] >= 0 and isinstance(e["binding"], str) and e["warp_instructions"] > 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"issue_slots"' in src and '"cpu_baseline"' in src and '"roofline"' in src and '"e2e"' in src


def test_variant_library_path_is_honoured(monkeypatch):
    """SYNSEG_LIB (tuning: a variant built with SYNSEG_BUILD_TAG) replaces the path of the product library; unset, the in-tree one loads."""
    import importlib
    from synapta_image_segmentation_b200 import _lib
    here = os.path.dirname(os.path.abspath(_lib.__file__))
    monkeypatch.setenv("SYNSEG_LIB", "/nonexistent/libsynseg_x.so")
    assert importlib.reload(_lib).LIB_PATH == "/nonexistent/libsynseg_x.so"
    monkeypatch.delenv("SYNSEG_LIB")
    assert importlib.reload(_lib).LIB_PATH == os.path.join(here, "libsynseg.so")
