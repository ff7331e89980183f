"""GPU parity tests of the device-side region stage (csrc/regions.cu), the grey ('L') page entry, the host entry points
that deliver validated regions, and the renderer-facing page slots -- all through the C ABI.

Checker: the host rules of geometry.py / detector.py (golden-tested against the imported reference in test_host_logic.py),
PIL's own grey conversion + numpy moments, and the cv2 chain."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from PIL import Image  # noqa: E402

from synapta_image_segmentation_b200 import _lib  # noqa: E402
from synapta_image_segmentation_b200.detector import DetectConfig, RasterRegionDetector  # noqa: E402
from synapta_image_segmentation_b200.ops import Context, PageSlots  # noqa: E402
from synapta_image_segmentation_b200.streaming import PageStreamer  # noqa: E402
from synapta_image_segmentation_b200.synth import page_shape, synth_page, synth_pages  # noqa: E402


def _key(r):
    b = r["bbox"]
    return (r["detection_method"], b.x0, b.y0, b.x1, b.y1, r["notes"], r.get("crop_px"), r.get("confidence"), r.get("validation"), r.get("variance"))


def _random_table(rng, n, w, h, mode):
    """A component table like connectedComponentsWithStats writes it: row 0 = background, then n boxes inside the page."""
    st = np.zeros((n + 1, 5), np.int32)
    st[0] = (0, 0, w, h, w * h)
    for k in range(1, n + 1):
        if mode == 0:       # specks and words: mostly small boxes, dense -> many clusters
            bw, bh = int(rng.integers(1, 90)), int(rng.integers(1, 40))
        elif mode == 1:     # a mix with figure-sized components
            bw, bh = (int(rng.integers(200, w // 2)), int(rng.integers(200, h // 2))) if rng.random() < 0.15 else (int(rng.integers(1, 120)), int(rng.integers(1, 120)))
        else:               # a lattice of equal boxes: gaps of exactly (60, 80) px -> distance exactly 100 at 72 DPI
            bw = bh = 20
        if mode == 2:
            x, y = 20 * int(rng.integers(0, (w - 20) // 20)), 20 * int(rng.integers(0, (h - 20) // 20))
        else:
            x, y = int(rng.integers(0, w - bw)), int(rng.integers(0, h - bh))
        st[k] = (x, y, bw, bh, max(1, int(bw * bh * rng.uniform(0.3, 1.0))))
    return st


@pytest.mark.parametrize("dpi", [300, 150, 72, 96])
def test_device_regions_equal_the_host_rules_on_random_tables(ctx, dpi):
    """synseg_regions_from_stats == detector.raster_regions (boxes, order, notes, crops) and PIL-grey numpy moments, for every
    page the device does not flag; flagged pages are exactly those the kernel must not decide (ambiguous distance)."""
    rng = np.random.default_rng(dpi)
    h, w = page_shape(dpi)
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=1100, max_regions=96), ctx=ctx)
    pw, ph = w * 72.0 / dpi, h * 72.0 / dpi
    B = 24
    tables = [_random_table(rng, int(rng.integers(0, 1000 if i % 4 else 60)), w, h, i % 3) for i in range(B)]
    tables[0] = _random_table(rng, 0, w, h, 0)                      # empty page
    page = synth_page(3, dpi, n_figures=2)[0]
    pages = torch.from_numpy(np.stack([page] * B)).cuda()
    stats = torch.zeros((B, 1100, 5), dtype=torch.int32)
    n = torch.zeros(B, dtype=torch.int32)
    for i, t in enumerate(tables):
        stats[i, :len(t)] = torch.from_numpy(t); n[i] = len(t)
    regions, n_regions, flags = ctx.regions_from_stats(n.cuda(), stats.cuda(), pages, dpi, pw, ph, max_regions=96)
    reg = Context.regions_view(regions.cpu()); n_regions = n_regions.cpu().numpy(); flags = flags.cpu().numpy()
    grey = np.array(Image.fromarray(page).convert("L")).astype(np.int64)
    flagged = 0
    for i, t in enumerate(tables):
        want = det.raster_regions(t, len(t), pw, ph)
        if flags[i]:
            flagged += 1
            assert flags[i] in (_lib.FLAG_AMBIGUOUS, _lib.FLAG_CAPACITY, _lib.FLAG_AMBIGUOUS | _lib.FLAG_CAPACITY)
            if flags[i] & _lib.FLAG_CAPACITY:
                assert len(want) > 96
            continue
        got = det.regions_from_table(reg[i], int(n_regions[i]), pw, ph)
        assert len(got) == len(want), (i, len(got), len(want))
        for g, v in zip(got, want):
            assert (g["detection_method"], g["notes"]) == (v["detection_method"], v["notes"])
            assert (g["bbox"].x0, g["bbox"].y0, g["bbox"].x1, g["bbox"].y1) == (v["bbox"].x0, v["bbox"].y0, v["bbox"].x1, v["bbox"].y1)   # bit-identical floats
            assert type(g["bbox"].x0) is type(v["bbox"].x0) and type(g["bbox"].y0) is type(v["bbox"].y0)      # the int 0 of max(0, .) too
            x, y, cw, chh = det._crop_px(v["bbox"], w, h)
            assert g["crop_px"] == (x, y, cw, chh)
            c = grey[y:y + chh, x:x + cw]
            assert g["_moments"] == (int(c.sum()), int((c * c).sum()))
    if dpi in (72,):
        assert flagged > 0          # lattice pages at 72 DPI hold pairs at exactly 100 pt: the device hands them to the host
    assert flagged < B


def _host_path(det, pages_np, pw, ph):
    """The round-1 flow: component tables -> host rules -> host-listed crop moments -> score -> keep -> sort."""
    t = torch.from_numpy(pages_np).cuda()
    n, stats, _ = det.detect_components(t)
    n = n.cpu().numpy(); stats = stats.cpu().numpy()
    h, w = pages_np.shape[1], pages_np.shape[2]
    out = []
    for i in range(pages_np.shape[0]):
        regs = det.raster_regions(stats[i], int(n[i]), pw, ph)
        page = t[i]
        det._host_moments(page, regs, w, h)
        det._score(regs, ph)
        kept = [r for r in regs if r["confidence"] >= det.cfg.keep_score]
        kept.sort(key=lambda r: (r["bbox"].y0, r["bbox"].x0))
        out.append(kept)
    return out


@pytest.mark.parametrize("dpi", [72, 150])
def test_detect_regions_batch_equals_the_host_path(ctx, dpi):
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    pages = synth_pages(8, dpi, base_seed=31)
    h, w = pages.shape[1], pages.shape[2]
    pw, ph = w * 72.0 / dpi, h * 72.0 / dpi
    got = det.detect_regions_batch(torch.from_numpy(pages).cuda())
    want = _host_path(det, pages, pw, ph)
    assert sum(len(g) for g in got) >= 4
    for g, v in zip(got, want):
        assert [_key(r) for r in g] == [_key(r) for r in v]
    # the variance is np.var of the PIL grey crop the reference scores (pdf_image_segmentation.py:2988-2989)
    for page, regs in zip(pages, got):
        for r in regs:
            x, y, cw, chh = r["crop_px"]
            gcrop = np.array(Image.fromarray(np.ascontiguousarray(page[y:y + chh, x:x + cw])).convert("L"))
            assert abs(r["variance"] - float(np.var(gcrop))) <= 1e-9 * max(1.0, r["variance"])


def test_label_overflow_is_retried_not_fatal(ctx):
    """A page with more components than max_labels (halftone dots) is re-run with a larger table; the batch survives."""
    dpi = 72
    h, w = page_shape(dpi)
    dots = np.full((h, w, 3), 255, np.uint8)
    dots[40:h - 40:60, 40:w - 40:60] = 0                      # ~100 isolated dots, far apart (k = 11 at 72 DPI)
    pages = np.stack([synth_page(1, dpi, n_figures=1)[0], dots, synth_page(2, dpi, n_figures=2)[0]])
    small = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=16), ctx=ctx)
    large = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=2048), ctx=ctx)
    t = torch.from_numpy(pages).cuda()
    tab = small.detect_tables(t, w * 72.0 / dpi, h * 72.0 / dpi)
    assert int(tab["n_labels"][1]) < 0 and int(tab["flags"][1]) == _lib.FLAG_LABELS and int(tab["n_regions"][1]) == 0
    got, want = small.detect_regions_batch(t), large.detect_regions_batch(t)
    for g, v in zip(got, want):
        assert [_key(r) for r in g] == [_key(r) for r in v]


def test_grey_pages_equal_rgb_pages(ctx):
    """channels = 1: an 'L' page gives the tables and regions of the same page handed over as RGB (cv2 and PIL grey of
    (g, g, g) are g), against the cv2 chain on the grey page."""
    from oracle import cv2_chain
    dpi = 150
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    rgb = synth_pages(5, dpi, base_seed=77)
    import cv2
    grey = np.stack([cv2.cvtColor(p, cv2.COLOR_RGB2GRAY) for p in rgb])
    rgb_of_grey = np.repeat(grey[..., None], 3, axis=3)
    n1, s1, c1 = det.detect_components(torch.from_numpy(grey).cuda())
    n2, s2, c2 = det.detect_components(torch.from_numpy(rgb_of_grey).cuda())
    n3, s3, _ = det.detect_components(torch.from_numpy(rgb).cuda())           # cv2 grey happens inside: same tables again
    assert torch.equal(n1, n2) and torch.equal(n1, n3)
    for j in range(5):
        m = int(n1[j])
        assert torch.equal(s1[j, :m], s2[j, :m]) and torch.equal(c1[j, :m], c2[j, :m]) and torch.equal(s1[j, :m], s3[j, :m])
    for j in range(2):                 # and the cv2 chain itself on the page the grey came from
        ref = cv2_chain.page_chain(rgb[j], dpi)
        assert int(n1[j]) == ref["n"] and np.array_equal(s1[j, :ref["n"]].cpu().numpy(), ref["stats"])
    # unaligned grey rows (width not a multiple of 16 -> tensor stride not 16-byte aligned) take the byte-load path
    odd = grey[:, :, :grey.shape[2] - 3].copy()
    na, sa, _ = det.detect_components(torch.from_numpy(odd).cuda())
    nb, sb, _ = det.detect_components(torch.from_numpy(np.repeat(odd[..., None], 3, axis=3)).cuda())
    assert torch.equal(na, nb) and all(torch.equal(sa[j, :int(na[j])], sb[j, :int(na[j])]) for j in range(5))
    ga = det.detect_regions_batch(torch.from_numpy(grey).cuda())
    gb = det.detect_regions_batch(torch.from_numpy(rgb_of_grey).cuda())
    for a, b in zip(ga, gb):
        assert [_key(r) for r in a] == [_key(r) for r in b]


@pytest.mark.parametrize("channels", [3, 1])
def test_streamer_regions_equal_detect_regions_batch(ctx, channels):
    """PageStreamer (synseg_detect_regions_host: staging ring, chunks, results in pinned memory) delivers exactly the
    validated regions of detect_regions_batch -- RGB and grey pages, ragged last batch and chunk."""
    dpi = 72
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    h, w = page_shape(dpi)
    import cv2
    batches = []
    for i, nb in enumerate((7, 7, 7, 3)):
        p = synth_pages(nb, dpi, base_seed=9, start=7 * i)
        if channels == 1:
            p = np.stack([cv2.cvtColor(q, cv2.COLOR_RGB2GRAY) for q in p])
        batches.append(torch.from_numpy(p).pin_memory())
    st = PageStreamer(det, 7, h, w, slots=2, chunk_pages=3, channels=channels)
    got, raw = {}, {}
    n_pages = st.run(iter(batches), on_result=lambda i, n, s: raw.__setitem__(i, (n.clone(), s.clone())),
                     on_regions=lambda i, regs: got.__setitem__(i, regs), page_base=100)
    assert n_pages == 24 and st.h2d_bytes == 24 * h * w * channels
    first = 0
    for i, hb in enumerate(batches):
        want = det.detect_regions_batch(hb.cuda(), page_nums=list(range(100 + first, 100 + first + hb.shape[0])))
        assert len(got[i]) == hb.shape[0]
        for a, b in zip(got[i], want):
            assert [_key(r) for r in a] == [_key(r) for r in b] and [r["page_num"] for r in a] == [r["page_num"] for r in b]
        n, stats, _ = det.detect_components(hb.cuda())
        assert torch.equal(raw[i][0], n.cpu())
        for j in range(hb.shape[0]):
            assert torch.equal(raw[i][1][j, :int(n[j])], stats[j, :int(n[j])].cpu())
        first += hb.shape[0]
    assert sum(len(r) for regs in got.values() for r in regs) >= 8


def test_streamer_recomputes_flagged_pages_on_the_host(ctx):
    """max_labels far too small for the pages: every page is flagged on the device and re-run; results still equal."""
    dpi = 72
    h, w = page_shape(dpi)
    tiny = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=4), ctx=ctx)
    full = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    hb = torch.from_numpy(synth_pages(4, dpi, base_seed=13)).pin_memory()
    got = {}
    PageStreamer(tiny, 4, h, w, slots=2, chunk_pages=2).run(iter([hb]), on_regions=lambda i, r: got.__setitem__(i, r))
    want = full.detect_regions_batch(hb.cuda())
    for a, b in zip(got[0], want):
        assert [_key(r) for r in a] == [_key(r) for r in b]


@pytest.mark.parametrize("channels", [1, 3])
def test_page_slots_renderer_handoff(ctx, channels):
    """synseg_page_slot_*: a renderer writes pages straight into the library's pinned slots; three slots in flight; the
    results equal the device entry point.  Misuse (submit without acquire, wait without submit) is an error, not a hang."""
    from synapta_image_segmentation_b200._lib import SynsegError
    dpi = 72
    h, w = page_shape(dpi)
    own = Context(0)                                   # the slots pin a geometry in the context's staging ring: use a private one
    try:
        det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=256, max_regions=32), ctx=own)
        bs, c, k = det.cfg.resolved()
        pw, ph = w * 72.0 / dpi, h * 72.0 / dpi
        slots = PageSlots(own, w, h, channels=channels, pages_per_slot=4, n_slots=3, max_labels=256, max_regions=32)
        assert slots.numa_node >= -1
        import cv2
        src = synth_pages(10, dpi, base_seed=41)
        if channels == 1:
            src = np.stack([cv2.cvtColor(q, cv2.COLOR_RGB2GRAY) for q in src])
        results, inflight = [], []
        for first in range(0, 10, 4):
            nb = min(4, 10 - first)
            s, buf = slots.acquire()
            buf[:nb] = src[first:first + nb]            # "rendering" into pinned memory
            slots.submit(s, nb, bs, c, k, dpi, pw, ph)
            inflight.append((s, first, nb))
            if len(inflight) == 3:
                s0, f0, n0 = inflight.pop(0)
                results.append((f0, n0, {kk: v.copy() for kk, v in slots.wait(s0).items()}))
        for s0, f0, n0 in inflight:
            results.append((f0, n0, {kk: v.copy() for kk, v in slots.wait(s0).items()}))
        want = det.detect_tables(torch.from_numpy(src).cuda(), pw, ph)
        wn, wr, wnr = want["n_labels"].cpu().numpy(), Context.regions_view(want["regions"].cpu()), want["n_regions"].cpu().numpy()
        ws = want["stats"].cpu().numpy()
        for f0, n0, res in results:
            assert np.array_equal(res["n_labels"], wn[f0:f0 + n0]) and np.array_equal(res["n_regions"], wnr[f0:f0 + n0])
            assert not res["flags"].any()
            for j in range(n0):
                assert np.array_equal(res["regions"][j, :res["n_regions"][j]], wr[f0 + j, :wnr[f0 + j]])
                assert np.array_equal(res["stats"][j, :res["n_labels"][j]], ws[f0 + j, :wn[f0 + j]])
        with pytest.raises(SynsegError):
            slots.submit(1, 1, bs, c, k, dpi, pw, ph)               # not acquired
        s, _ = slots.acquire()
        with pytest.raises(SynsegError):
            slots.wait(s)                                           # not submitted
        slots.close()
    finally:
        own.close()


def test_calls_on_different_streams_are_ordered(ctx):
    """All calls of a context share one scratch arena; a call arriving on another stream waits for the previous one
    (CallGuard in csrc/internal.cuh), so interleaving streams cannot corrupt results."""
    dpi = 150
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    pages = torch.from_numpy(synth_pages(6, dpi, base_seed=3)).cuda()
    want = det.detect_components(pages)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(6):
        with torch.cuda.stream(s1 if rep % 2 == 0 else s2):
            outs.append(det.detect_components(pages))
    torch.cuda.synchronize()
    for n, st, _ in outs:
        assert torch.equal(n, want[0])
        for j in range(6):
            assert torch.equal(st[j, :int(n[j])], want[1][j, :int(n[j])])


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_contexts_on_two_devices(ctx):
    """per-device kernel attributes and device handling -- a second context on GPU 1 under current device 0 runs the
    300-DPI path (k = 41 folds to 81: the van Herk column kernel needs > 48 KB of dynamic shared memory) and leaves the
    caller's current device alone."""
    torch.cuda.set_device(0)
    other = Context(1)
    try:
        assert torch.cuda.current_device() == 0
        det0 = RasterRegionDetector(DetectConfig(dpi=300, max_labels=512), ctx=ctx)
        det1 = RasterRegionDetector(DetectConfig(dpi=300, max_labels=512), ctx=other)
        pages = synth_pages(2, 300, base_seed=5)
        a = det0.detect_components(torch.from_numpy(pages).cuda(0))
        b = det1.detect_components(torch.from_numpy(pages).cuda(1))
        assert torch.cuda.current_device() == 0
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        assert torch.equal(a[0].cpu(), b[0].cpu())
        for j in range(2):
            assert torch.equal(a[1][j, :int(a[0][j])].cpu(), b[1][j, :int(a[0][j])].cpu())
        with pytest.raises(ValueError):
            det1.detect_components(torch.from_numpy(pages).cuda(0))          # tensor on the wrong device
    finally:
        other.close()


def test_two_pass_priors_on_the_gpu(ctx):
    """Pass 2 of _extract_images_from_page through the detector: validated raster regions are candidates resolved, in detection
    order, against the caption-based priors by find_conflicting / resolve_conflict (golden-tested in test_host_logic.py)."""
    from synapta_image_segmentation_b200 import geometry as G
    from synapta_image_segmentation_b200.datamodel import BoundingBox
    dpi = 150
    det = RasterRegionDetector(DetectConfig(dpi=dpi), ctx=ctx)
    page = synth_page(0, dpi, n_figures=2)[0]
    h, w = page.shape[:2]
    base = det.detect_regions(page, 0)
    assert len(base) >= 2
    # the validated raster candidates in DETECTION order (the order pass 2 walks them in)
    t = det.detect_tables(torch.from_numpy(page).cuda()[None], 612.0, 792.0)
    cands = det.regions_from_table(Context.regions_view(t["regions"].cpu())[0], int(t["n_regions"][0]), 612.0, 792.0)
    det._score(cands, 792.0)
    cands = [r for r in cands if r["confidence"] >= 0.5]
    assert sorted(_key(r) for r in cands) == sorted(_key(r) for r in base)
    tgt = min(base, key=lambda r: r["bbox"].area())["bbox"]
    priors = {
        "captioned, larger": {"bbox": BoundingBox(tgt.x0 - 5, tgt.y0 - 5, tgt.x1 + 5, tgt.y1 + 40, 612.0, 792.0), "caption": "Figure 1.1 Returns",
                              "notes": "Caption: Figure 1.1"},
        "no caption, same box": {"bbox": BoundingBox(tgt.x0, tgt.y0, tgt.x1, tgt.y1, 612.0, 792.0), "caption": None},
        "far away": {"bbox": BoundingBox(2, 2, 40, 40, 612.0, 792.0), "caption": "Exhibit 9"},
    }
    seen = set()
    for name, prior in priors.items():
        out = det.detect_regions(page, 0, priors=[prior])
        scored = dict(prior, detection_method="caption_based", confidence=0.9, caption=prior.get("caption"))
        want = G.resolve_page_conflicts([scored], [dict(c) for c in cands])
        want.sort(key=lambda r: (r["bbox"].y0, r["bbox"].x0))
        assert [(r["detection_method"], r["bbox"]) for r in out] == [(r["detection_method"], r["bbox"]) for r in want], name
        for r in out:
            if r["detection_method"] == "caption_based":
                assert r["confidence"] == 0.9 and r["caption"] == prior["caption"] and "variance" in r and "crop_px" in r
        seen.add(("caption_based" in [r["detection_method"] for r in out], any("conflict_resolution" in r for r in out)))
    assert any(stays for stays, _ in seen) and any(not stays for stays, _ in seen)      # a prior that stays, and one a raster region replaced
    far = det.detect_regions(page, 0, priors=[priors["far away"]])
    alone = det.detect_regions(page, 0, priors=[])                      # pass 2 among the raster candidates only
    assert len(far) == len(alone) + 1 and len(alone) <= len(base)
    assert [r["bbox"] for r in alone] == [r["bbox"] for r in sorted(G.resolve_page_conflicts([], [dict(c) for c in cands]), key=lambda r: (r["bbox"].y0, r["bbox"].x0))]
    # the other rule (dead code in the reference, kept selectable): _detect_visual_regions' duplicate test
    out3 = det.detect_regions(page, 0, priors=[priors["captioned, larger"]], prior_rule="visual_regions")
    assert [r["detection_method"] for r in out3].count("caption_based") == 1
    segs = det.extract_segments(page, 0, "textbook_001", priors=[priors["far away"]])
    capseg = [s for s in segs if s.extraction_method == "caption_based"][0]
    assert capseg.caption_text == "Exhibit 9" and capseg.notes == "" and capseg.confidence == 0.9 and capseg.page_no == 1


def test_hints_on_resident_regions_equal_hints_on_crops(ctx):
    """synseg_hints_rois / synseg_colors_rois read the detector's regions in place on the pages in HBM; same numbers as the crops
    cut out, packed and uploaded (hints_batch), which the goldens pin to the reference."""
    from synapta_image_segmentation_b200.hints import FeatureHints
    dpi = 150
    det = RasterRegionDetector(DetectConfig(dpi=dpi), ctx=ctx)
    pages = synth_pages(6, dpi, base_seed=17)
    t = torch.from_numpy(pages).cuda()
    regions = det.detect_regions_batch(t)
    got = FeatureHints.hints_regions(t, regions)
    crops = [Image.fromarray(np.ascontiguousarray(pages[i][y:y + h, x:x + w])) for i, regs in enumerate(regions) for (x, y, w, h) in [r["crop_px"] for r in regs]]
    want = FeatureHints.hints_batch(crops)
    flat = [hd for page in got for hd in page]
    assert len(flat) == len(want) >= 6
    assert flat == want
    # a region that ends on the very last pixel of the last page, and grey pages (no colours)
    h, w = pages.shape[1], pages.shape[2]
    edge = [[] for _ in range(5)] + [[{"crop_px": (w - 333, h - 200, 333, 200)}]]
    e = FeatureHints.hints_regions(t, edge)[5][0]
    assert e == FeatureHints.hints_batch([Image.fromarray(np.ascontiguousarray(pages[5][h - 200:, w - 333:]))])[0]
    import cv2
    grey = np.stack([cv2.cvtColor(p, cv2.COLOR_RGB2GRAY) for p in pages])
    g = FeatureHints.hints_regions(torch.from_numpy(grey).cuda(), regions)
    gw = FeatureHints.hints_batch([Image.fromarray(np.ascontiguousarray(grey[i][y:y + hh, x:x + ww])) for i, regs in enumerate(regions)
                                   for (x, y, ww, hh) in [r["crop_px"] for r in regs]])
    assert [{k: v for k, v in d.items()} for page in g for d in page] == [{k: v for k, v in d.items() if k not in ("dominant_colors", "color_weights")} for d in gw]


def test_dedup_exchange_equals_sorted_all_pairs(ctx):
    """synseg_dedup_exchange at world = 1 (pack -> compact -> rank + all-pairs -> scatter): keys ascending, keep flags equal to the
    plain all-pairs rule on the sorted list; entries past *count are ignored; order of the input does not matter."""
    from synapta_image_segmentation_b200.dedup import DedupExchange, survivors_digest
    rng = np.random.default_rng(8)
    n, cap = 3000, 4096
    base = rng.integers(0, 2 ** 63 - 1, 600, dtype=np.int64)
    hashes = base[rng.integers(0, 600, n)] ^ (np.int64(1) << rng.integers(0, 63, n)) * (rng.random(n) < 0.5)      # near-duplicates (1 bit apart)
    keys = rng.permutation(n * 7)[:n].astype(np.int64) << 16
    ks = np.argsort(keys)
    hk, kk = hashes[ks], keys[ks]
    x = (hk[:, None] ^ hk[None, :]).view(np.uint64)
    pop = np.zeros(x.shape, np.int32)
    for b in range(64):
        pop += ((x >> np.uint64(b)) & np.uint64(1)).astype(np.int32)
    want_keep = ~((pop <= 4) & (np.arange(n)[None, :] < np.arange(n)[:, None])).any(1)
    for trial in range(2):
        perm = rng.permutation(n)
        hb = torch.zeros(cap, dtype=torch.int64); kb = torch.zeros(cap, dtype=torch.int64)
        hb[:n] = torch.from_numpy(hashes[perm]); kb[:n] = torch.from_numpy(keys[perm])
        hb[n:] = 12345; kb[n:] = 1                                      # garbage past count must be ignored
        ex = DedupExchange(ctx, cap, 1, 4)
        ex.run(hb.cuda(), kb.cuda(), torch.tensor([n], dtype=torch.int32, device="cuda"))
        k, keep = ex.result()
        assert k.numpy().tolist() == kk.tolist()
        assert keep.numpy().astype(bool).tolist() == want_keep.tolist()
        assert len(survivors_digest(k, keep)) == 16
    ex.run(hb.cuda(), kb.cuda(), torch.zeros(1, dtype=torch.int32, device="cuda"))
    k, keep = ex.result()
    assert k.numel() == 0


def test_fused_morphology_and_separate_front_end_variants(ctx, monkeypatch):
    """The opt-in one-kernel dilate + close (SYNSEG_FUSED_MORPH=1, csrc/morph_fused.cu) and the non-TMA front end (SYNSEG_NO_TMA=1:
    rgb2gray + canny_classes) give the tables of the default path (TMA-fed fused front end, four morphology passes) on odd sizes,
    several element sizes (even k included) and a size whose last 16-byte unit is partial."""
    from oracle import cv2_chain
    import cv2
    rng = np.random.default_rng(12)
    for (h, w, bs, k) in [(333, 517, 15, 7), (200, 1100, 25, 12), (792, 612, 13, 11), (97, 131, 15, 5), (1650, 1275, 25, 21)]:
        page = synth_page(int(rng.integers(100)), 150 if h == 1650 else 72, n_figures=2)[0][:h, :w]
        if page.shape[0] < h or page.shape[1] < w:
            page = np.ascontiguousarray(np.pad(page, ((0, h - page.shape[0]), (0, w - page.shape[1]), (0, 0)), constant_values=255))
        page = np.ascontiguousarray(page)
        t = torch.from_numpy(np.stack([page, page[::-1].copy(), page[:, ::-1].copy()])).cuda()
        results = []
        for env in ({}, {"SYNSEG_FUSED_MORPH": "1"}, {"SYNSEG_NO_TMA": "1"}, {"SYNSEG_FUSED_MORPH": "1", "SYNSEG_NO_TMA": "1"}):
            for kk in ("SYNSEG_FUSED_MORPH", "SYNSEG_NO_TMA"):
                monkeypatch.delenv(kk, raising=False)
            for kk, vv in env.items():
                monkeypatch.setenv(kk, vv)
            l0 = ctx.launches
            gray = torch.empty((3, h, (w + 15) // 16 * 16), dtype=torch.uint8, device="cuda")[:, :, :w]
            n, st, ce = ctx.detect_pages(t, bs, 10, k, max_labels=2048, gray_out=gray)
            torch.cuda.synchronize()
            results.append((n.cpu(), st.cpu(), ce.cpu(), gray.cpu(), ctx.launches - l0))
        for kk in ("SYNSEG_FUSED_MORPH", "SYNSEG_NO_TMA"):
            monkeypatch.delenv(kk, raising=False)
        n0, st0, ce0, g0, launches0 = results[0]
        assert np.array_equal(g0[0].numpy(), cv2.cvtColor(page, cv2.COLOR_RGB2GRAY))      # the TMA kernel writes cv2's grey plane
        for n, st, ce, g, launches in results[1:]:
            assert torch.equal(n, n0) and torch.equal(g, g0)
            for j in range(3):
                m = int(n0[j])
                assert torch.equal(st[j, :m], st0[j, :m]) and torch.equal(ce[j, :m], ce0[j, :m])
        assert results[1][4] == launches0 - 3 and results[2][4] == launches0 + 1      # 4 morphology passes -> 1 kernel; 1 front kernel -> 2
        # and the cv2 chain itself (dilate k, close k)
        g = cv2.cvtColor(page, cv2.COLOR_RGB2GRAY)
        thr = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, bs, 10)
        ink = cv2.bitwise_or(thr, cv2.Canny(g, 50, 150))
        se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        closed = cv2.morphologyEx(cv2.dilate(ink, se), cv2.MORPH_CLOSE, se)
        nn, _, ss, cc = cv2.connectedComponentsWithStats(closed, 8, cv2.CV_32S)
        assert int(n0[0]) == nn and np.array_equal(st0[0, :nn].numpy(), ss) and np.array_equal(ce0[0, :nn].numpy(), cc)


def test_cuda_graph_capture_of_the_detection_chain(ctx):
    """Context.capture: the ~45 launches of a detection step (two chains on two streams, TMA kernel, hysteresis sweeps, union-find)
    replayed as ONE CUDA graph give the tables of the direct call -- also after the input pages change in place."""
    dpi = 150
    det = RasterRegionDetector(DetectConfig(dpi=dpi, max_labels=512), ctx=ctx)
    a = synth_pages(6, dpi, base_seed=51)
    b = synth_pages(6, dpi, base_seed=52)
    pages = torch.from_numpy(a).cuda()
    out = (torch.empty(6, dtype=torch.int32, device="cuda"), torch.empty((6, 512, 5), dtype=torch.int32, device="cuda"),
           torch.empty((6, 512, 2), dtype=torch.float64, device="cuda"))
    replay = ctx.capture(lambda: det.detect_components(pages, out=out))
    for src in (a, b, a):
        pages.copy_(torch.from_numpy(src).cuda())
        for t_ in out:
            t_.zero_()
        l0 = ctx.launches
        replay()
        torch.cuda.synchronize()
        assert ctx.launches == l0                           # nothing launched kernel by kernel
        got = [t_.clone() for t_ in out]
        want = det.detect_components(torch.from_numpy(src).cuda())
        assert torch.equal(got[0], want[0])
        for j in range(6):
            m = int(want[0][j])
            assert m > 1 and torch.equal(got[1][j, :m], want[1][j, :m]) and torch.equal(got[2][j, :m], want[2][j, :m])
