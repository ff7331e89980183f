"""Generates tests/golden/* by importing the reference in THIS container (never on the GPU box).

    python tests/golden/make_golden.py

* copies 64 real crops from /root/reference/investments_segmented (inputs only, data not source; among them
  the 1191x1500 Appendix-D crop and 40+ crops wider than one 480-column strip) plus synthetic crops, and records
  what the reference's own functions return on them: OCRProcessor._detect_grid / _count_arrows / _detect_shapes /
  _estimate_data_points / _detect_chart_subtype (no OCR and four OCR texts) / np.var / mask counts
  (pdf_image_segmentation.py:1320-1461, 1546-1617, 1753-1810);
* records the same deterministic quantities for ALL 591 crops of the shipped run in reference_corpus.json and copies
  the crops to tests/golden/_a2/ (git-ignored: 38 MB; they travel to the GPU box with the tree, the JSON is committed);
* records the pure-geometry known answers of SURVEY.md Appendix D by calling
  _calculate_overlap_ratio, _overlaps_with_existing, _drawing_distance, _cluster_drawings,
  _detect_by_drawings and _validate_embedded_image on an uninitialised pipeline instance.
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import shutil
import sys

import cv2
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from synapta_image_segmentation_b200.synth import render_figure  # noqa: E402

ref = ref_import.load()
A2 = os.path.join(ref_import.REFERENCE_DIR, "investments_segmented")
BASE = ["textbook_001_p020_2b1be7d6.png", "textbook_001_p022_3c6ac748.png", "textbook_001_p022_f96abaf9.png",
        "textbook_001_p023_e2cf5878.png",
        # a spread over the book (RGB and grey, 245..548 rows), each < 40 KB
        "textbook_001_p169_513ec97a.png", "textbook_001_p183_af3dbb7a.png", "textbook_001_p207_364b0248.png",
        "textbook_001_p269_7d5d6fda.png", "textbook_001_p405_1ee5f3bc.png", "textbook_001_p457_ee651881.png",
        "textbook_001_p510_5645a3d3.png", "textbook_001_p553_a84c5fb1.png", "textbook_001_p712_8719d410.png",
        "textbook_001_p826_62601fad.png", "textbook_001_p971_84e35f5e.png", "textbook_001_p973_d3eae19d.png",
        "textbook_001_p988_5b5815f5.png",
        # SURVEY.md Appendix D known-answer crop (1191 x 1500: arrows 20, rectangles 89) and the next largest ones
        "textbook_001_p000_ab84f0ff.png", "textbook_001_p948_749fb1c3.png", "textbook_001_p179_aa2281bf.png"]
OCR_TEXTS = ["Figure 3.2 Bar chart of annual returns by asset class", "A line graph of the yield curve: rates (%) against maturity",
             "Pie chart: portfolio weights", "Open High Low Close prices; scatter of risk and return"]


def real_selection():
    """BASE + every 13th file of the sorted corpus below 160 KB (deterministic): 64 crops."""
    names = sorted(f for f in os.listdir(A2) if f.endswith(".png"))
    out = list(BASE)
    for f in names[::13]:
        if f not in out and os.path.getsize(os.path.join(A2, f)) < 160_000 and len(out) < 64:
            out.append(f)
    for f in names[5::29]:
        if f not in out and os.path.getsize(os.path.join(A2, f)) < 160_000 and len(out) < 64:
            out.append(f)
    return out


def quiet(fn, *a):
    """The reference's _detect_chart_subtype prints DEBUG lines (pdf_image_segmentation.py:1379-1451)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def helper_record(path, colours=True):
    img = Image.open(path)
    g = np.array(img.convert("L"))
    edges = cv2.Canny(g, 50, 150)
    hk = cv2.getStructuringElement(cv2.MORPH_RECT, (25, 1)); vk = cv2.getStructuringElement(cv2.MORPH_RECT, (1, 25))
    h_count = int(np.sum(cv2.morphologyEx(edges, cv2.MORPH_OPEN, hk, iterations=2) > 0))
    v_count = int(np.sum(cv2.morphologyEx(edges, cv2.MORPH_OPEN, vk, iterations=2) > 0))
    rec = dict(size=list(img.size), mode=img.mode,
               detect_grid=bool(ref.OCRProcessor._detect_grid(img)),
               count_arrows=int(ref.OCRProcessor._count_arrows(img)),
               detect_shapes={k: int(v) for k, v in ref.OCRProcessor._detect_shapes(img).items()},
               estimate_data_points=int(ref.OCRProcessor._estimate_data_points(img)),
               connections=len(ref.OCRProcessor._extract_connections(img)),
               image_subtype=ref.OCRProcessor._detect_image_subtype(img, None),
               chart_subtype=quiet(ref.OCRProcessor._detect_chart_subtype, img, None),
               h_count=h_count, v_count=v_count, edge_px=int(np.sum(edges > 0)), variance=float(np.var(g)))
    assert rec["detect_grid"] == (h_count > 300 and v_count > 300)
    if not colours:
        if img.mode == "RGB":
            hsv = cv2.cvtColor(np.array(img), cv2.COLOR_RGB2HSV)
            rec["mask_px"] = int(((hsv[:, :, 1] > 30) & (hsv[:, :, 2] > 40) & (hsv[:, :, 2] < 240)).sum())
        else:
            rec["mask_px"] = 0
        return rec
    rec["chart_subtype_texts"] = [quiet(ref.OCRProcessor._detect_chart_subtype, img, ref.OCRResult(raw_text=t)) for t in OCR_TEXTS]
    if img.mode == "RGB":
        a = np.array(img)
        hsv = cv2.cvtColor(a, cv2.COLOR_RGB2HSV)
        rec["mask_px"] = int(((hsv[:, :, 1] > 30) & (hsv[:, :, 2] > 40) & (hsv[:, :, 2] < 240)).sum())
        np.random.seed(7)
        rec["dominant_colors_seed7"] = ref.OCRProcessor._extract_dominant_colors(img)
    else:
        rec["mask_px"] = 0
        rec["dominant_colors_seed7"] = ref.OCRProcessor._extract_dominant_colors(img)
    return rec


def corpus_record(name):
    return name, helper_record(os.path.join(A2, name), colours=False)


def main():
    crops = {}
    for name in real_selection():
        shutil.copy(os.path.join(A2, name), os.path.join(HERE, name))
        crops[name] = helper_record(os.path.join(HERE, name))
    # the whole shipped run ("parity on the 591 real crops")
    names = sorted(f for f in os.listdir(A2) if f.endswith(".png"))
    os.makedirs(os.path.join(HERE, "_a2"), exist_ok=True)
    for f in names:
        if not os.path.exists(os.path.join(HERE, "_a2", f)):
            shutil.copy(os.path.join(A2, f), os.path.join(HERE, "_a2", f))
    with mp.Pool(os.cpu_count()) as pool:
        corpus = dict(pool.map(corpus_record, names, chunksize=8))
    json.dump(dict(crops=corpus, ocr_texts=OCR_TEXTS, grid_true=sum(r["detect_grid"] for r in corpus.values())),
              open(os.path.join(HERE, "reference_corpus.json"), "w"), indent=0)
    for i in range(4):       # synthetic crops at 150 DPI
        fig = render_figure([99, i], 150, 500, 700)
        name = f"synth_fig_{i}.png"
        Image.fromarray(fig).save(os.path.join(HERE, name))
        crops[name] = helper_record(os.path.join(HERE, name))

    pl = ref_import.pipeline_instance(ref)
    BB = ref.BoundingBox
    geo = {}
    geo["overlap_ratio"] = pl._calculate_overlap_ratio(BB(0, 0, 10, 10, 612, 792), BB(5, 5, 20, 20, 612, 792))
    ex = [{"bbox": BB(0, 0, 100, 100, 612, 792)}]
    geo["overlaps_existing"] = [bool(pl._overlaps_with_existing(BB(50, 0, 150, 100, 612, 792), ex)),
                                bool(pl._overlaps_with_existing(BB(49, 0, 149, 100, 612, 792), ex))]
    geo["drawing_distance"] = [pl._drawing_distance([0, 0, 10, 10], [10, 10, 20, 20]), pl._drawing_distance([0, 0, 10, 10], [13, 14, 20, 20])]
    rects5 = [{"rect": [x, 0, x + 10, 10]} for x in (0, 60, 150, 240, 330)]
    geo["cluster_5"] = [[d["rect"] for d in c] for c in pl._cluster_drawings(rects5, None)]
    rng = np.random.default_rng(5)
    rects = []
    for _ in range(40):
        x, y = float(rng.uniform(0, 500)), float(rng.uniform(0, 700))
        rects.append([x, y, x + float(rng.uniform(2, 60)), y + float(rng.uniform(2, 60))])
    geo["cluster_random_in"] = rects
    geo["cluster_random_out"] = [[d["rect"] for d in c] for c in pl._cluster_drawings([{"rect": r} for r in rects], None)]

    class FakeRect:
        width, height = 612.0, 792.0

    class FakePage:
        rect = FakeRect()

        def __init__(self, drawings):
            self._d = drawings

        def get_drawings(self):
            return self._d

    regs = pl._detect_by_drawings(FakePage([{"rect": [100 + 5 * i, 100, 110 + 5 * i, 300]} for i in range(6)]), FakeRect())
    geo["detect_by_drawings"] = [dict(bbox=[r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1], caption=r["caption"],
                                      detection_method=r["detection_method"], notes=r["notes"]) for r in regs]
    regs = pl._detect_by_drawings(FakePage([{"rect": r} for r in rects]), FakeRect())
    geo["detect_by_drawings_random"] = [dict(bbox=[r["bbox"].x0, r["bbox"].y0, r["bbox"].x1, r["bbox"].y1], notes=r["notes"]) for r in regs]

    pl._find_caption_near_bbox = lambda page, bbox: None
    noise = Image.fromarray(np.random.default_rng(1).integers(0, 256, (300, 400, 3), dtype=np.uint8))
    flat = Image.fromarray(np.full((300, 300), 128, np.uint8))
    small = Image.fromarray(np.random.default_rng(2).integers(0, 256, (40, 300, 3), dtype=np.uint8))
    val = []
    for nm, im, bb in [("noise", noise, (100, 200, 300, 350)), ("noise", noise, (100, 200, 150, 250)), ("flat", flat, (100, 20, 300, 350)),
                       ("small", small, (100, 200, 400, 240)), ("noise", noise, (100, 750, 300, 790)), ("flat", flat, (10, 300, 600, 320))]:
        score, notes = pl._validate_embedded_image(im, BB(*bb, 612, 792), FakePage([]))
        val.append(dict(image=nm, bbox=list(bb), score=score, notes=notes))
    geo["validate"] = val

    class Seg:
        def __init__(self, bbox, caption_text=None, extraction_method="embedded_image", confidence=1.0, image_path=None):
            self.bbox, self.caption_text, self.extraction_method, self.confidence, self.image_path = bbox, caption_text, extraction_method, confidence, image_path
    npath = os.path.join(HERE, "_tmp_noise.png"); noise.save(npath)
    fpath = os.path.join(HERE, "_tmp_flat.png"); flat.save(fpath)
    rc = []
    for emb_bb, cap_bb, cap_text, conf, path in [((0, 0, 100, 100), (0, 0, 100, 130), "Figure 1", 0.9, npath),
                                                 ((0, 0, 100, 100), (0, 0, 100, 100), None, 0.9, npath),
                                                 ((0, 0, 200, 200), (0, 0, 100, 100), None, 0.5, fpath),
                                                 ((0, 0, 100, 100), (0, 0, 100, 125), None, 0.71, fpath)]:
        d, why = pl._resolve_conflict(Seg(BB(*emb_bb, 612, 792), confidence=conf, image_path=path),
                                      Seg(BB(*cap_bb, 612, 792), caption_text=cap_text, extraction_method="rendered_region"), FakePage([]))
        rc.append(dict(emb=list(emb_bb), cap=list(cap_bb), caption=cap_text, confidence=conf, image=os.path.basename(path)[5:-4], decision=d, reasons=why))
    geo["resolve_conflict"] = rc

    # Pass 2 of _extract_images_from_page (pdf_image_segmentation.py:2822-2847): every validated candidate is tested against the
    # segments kept so far with the reference's own _find_conflicting_segment / _resolve_conflict (the loop statements
    # around those two calls are restated here; `page.get_drawings()` feeds factor 4).
    def pass2(captions, candidates, drawings):
        segments = [Seg(BB(*c["bbox"], 612, 792), caption_text=c["caption"], extraction_method="caption_based", confidence=0.9) for c in captions]
        for c in candidates:
            cand = Seg(BB(*c["bbox"], 612, 792), confidence=c["confidence"], image_path=npath if c["image"] == "noise" else fpath)
            conflict = pl._find_conflicting_segment(cand, segments)
            if conflict:
                decision, _ = pl._resolve_conflict(cand, conflict, FakePage([{"rect": FakeDrawRect(*d)} for d in drawings]))
                if decision == "keep_embedded":
                    segments.remove(conflict)
                    segments.append(cand)
            else:
                segments.append(cand)
        return [dict(method=s_.extraction_method, bbox=[s_.bbox.x0, s_.bbox.y0, s_.bbox.x1, s_.bbox.y1]) for s_ in segments]

    class FakeDrawRect:
        def __init__(self, x0, y0, x1, y1):
            self.x0, self.y0, self.x1, self.y1 = x0, y0, x1, y1

    ref.fitz.Rect = FakeDrawRect          # `drawing.get('rect', fitz.Rect(0, 0, 0, 0))` builds the default eagerly (:3083)
    scenarios = [
        dict(captions=[dict(bbox=[50, 100, 300, 330], caption="Figure 2.1 Returns")],
             candidates=[dict(bbox=[60, 110, 290, 300], confidence=0.9, image="noise"),      # photo inside a captioned region: caption wins 5:3
                         dict(bbox=[320, 400, 560, 600], confidence=0.8, image="flat")],     # no conflict: added
             drawings=[]),
        dict(captions=[dict(bbox=[50, 100, 300, 300], caption=None)],
             candidates=[dict(bbox=[40, 90, 320, 330], confidence=0.9, image="noise")],      # no caption text, photo-like: embedded replaces it
             drawings=[]),
        dict(captions=[dict(bbox=[50, 100, 300, 300], caption=None), dict(bbox=[50, 400, 300, 600], caption="Fig. 3 Flow")],
             candidates=[dict(bbox=[60, 110, 290, 290], confidence=0.6, image="flat"),       # 0:0 -> keep_embedded (ties go to the embedded image)
                         dict(bbox=[70, 120, 280, 280], confidence=0.9, image="noise"),      # now conflicts with the candidate kept before it
                         dict(bbox=[60, 410, 290, 590], confidence=0.95, image="noise")],    # caption 3 vs embedded 2+1: tie -> keep_embedded
             drawings=[]),
        dict(captions=[dict(bbox=[50, 100, 300, 300], caption=None)],
             candidates=[dict(bbox=[60, 110, 290, 290], confidence=0.75, image="flat")],     # 12 drawings inside the caption region: caption 2 vs 1
             drawings=[[60 + 5 * i, 120, 70 + 5 * i, 130] for i in range(12)]),
    ]
    geo["pass2"] = [dict(s_, result=pass2(s_["captions"], s_["candidates"], s_["drawings"])) for s_ in scenarios]
    os.remove(npath); os.remove(fpath)

    json.dump(dict(crops=crops, ocr_texts=OCR_TEXTS, versions=dict(cv2=cv2.__version__, numpy=np.__version__)), open(os.path.join(HERE, "reference_helpers.json"), "w"), indent=1)
    json.dump(geo, open(os.path.join(HERE, "reference_geometry.json"), "w"), indent=1)
    print("wrote", len(crops), "crop records and geometry vectors")


if __name__ == "__main__":
    main()
