"""Pins the CPU oracle (oracle/synseg_oracle.c) against the live third-party primitives the reference
calls (cv2 4.13 / PIL 12 / numpy 2.3 -- un-vendored, unpinned by the reference) and against golden
vectors generated from the imported reference (tests/golden, made by tests/golden/make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest

import imgs
import oracle
from oracle import cv2_chain

cv2 = pytest.importorskip("cv2")
from PIL import Image  # noqa: E402

SIZES = [(1, 1), (2, 3), (7, 5), (31, 33), (64, 64), (97, 131), (240, 317)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gray_exhaustive():
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    rgb = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))
    assert np.array_equal(oracle.rgb2gray_cv(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(oracle.rgb2gray_pil(rgb), np.array(Image.fromarray(rgb, "RGB").convert("L")))
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    s, v = oracle.hsv_sv(rgb)
    assert np.array_equal(s, hsv[:, :, 1]) and np.array_equal(v, hsv[:, :, 2])
    m, n = oracle.hsv_mask(rgb)
    want = cv2_chain.hsv_mask(rgb)
    assert np.array_equal(m.astype(bool), want) and n == int(want.sum())


@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("bs,c,inv", [(3, 2, True), (15, 10, True), (25, 10, True), (51, 10, True), (31, 5, False), (101, 7, True)])
def test_adaptive(hw, bs, c, inv):
    g = imgs.blurred_noise(*hw, seed=5)
    want = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV if inv else cv2.THRESH_BINARY, bs, c)
    assert np.array_equal(oracle.adaptive_mean(g, bs, c, inv), want)


@pytest.mark.parametrize("hw", SIZES + [(513, 770)])
def test_canny(hw):
    h, w = hw
    for g in (imgs.blurred_noise(h, w, 21, passes=2), np.ascontiguousarray(imgs.shapes(max(h, 8), max(w, 8), 22)[:h, :w])):
        assert np.array_equal(oracle.canny(g, 50, 150), cv2.Canny(g, 50, 150))


def test_canny_thread_invariant():
    g = imgs.shapes(400, 600, 3)
    n = cv2.getNumThreads()
    try:
        cv2.setNumThreads(1); a = cv2.Canny(g, 50, 150)
        cv2.setNumThreads(8); b = cv2.Canny(g, 50, 150)
    finally:
        cv2.setNumThreads(n)
    assert np.array_equal(a, b) and np.array_equal(a, oracle.canny(g))


@pytest.mark.parametrize("kw,kh", [(3, 3), (1, 25), (25, 1), (21, 21), (20, 1), (1, 30), (4, 6)])
def test_morphology(kw, kh):
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
    for src in (imgs.random_mask(97, 131, 31, 0.7), imgs.blurred_noise(60, 85, 41, passes=1)):
        assert np.array_equal(oracle.morph_rect(src, 1, kw, kh), cv2.dilate(src, se))
        assert np.array_equal(oracle.morph_rect(src, 0, kw, kh), cv2.erode(src, se))
        for mop, cvop in ((2, cv2.MORPH_OPEN), (3, cv2.MORPH_CLOSE)):
            for it in (1, 2):
                assert np.array_equal(oracle.morphology_ex(src, mop, kw, kh, it), cv2.morphologyEx(src, cvop, se, iterations=it))


def test_morphology_iterations_fold():
    """iterations=2 of a k-rect == one pass with (2k-1) and anchor 2*(k//2)."""
    src = imgs.random_mask(80, 90, 1, 0.6)
    for k in (20, 25, 30):
        two = oracle.morph_rect(oracle.morph_rect(src, 0, k, 1), 0, k, 1)
        one = oracle.morph_rect(src, 0, 2 * k - 1, 1, 2 * (k // 2), 0)
        assert np.array_equal(two, one)


@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("density", [0.05, 0.3, 0.5, 0.62, 0.9])
def test_ccl(hw, density):
    m = imgs.random_mask(*hw, seed=int(density * 100), density=density)
    n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(m, 8, cv2.CV_32S)
    n, lab, st, ce = oracle.ccl8_stats(m)
    assert n == n_w and np.array_equal(lab, lab_w) and np.array_equal(st, st_w)
    assert np.array_equal(ce, ce_w, equal_nan=True)


def test_ccl_adversarial_and_thread_invariance():
    cases = [imgs.spiral_mask(201, 333), imgs.checkerboard(64, 65), np.zeros((17, 19), np.uint8), np.full((17, 19), 255, np.uint8)]
    m = np.zeros((6, 14), np.uint8)
    for (r, c) in [(1, 0), (0, 10), (4, 3), (3, 12)]:
        m[r, c] = 255
    cases.append(m)
    nthreads = cv2.getNumThreads()
    for m in cases:
        n, lab, st, ce = oracle.ccl8_stats(m)
        for nt in (1, 8):
            cv2.setNumThreads(nt)
            n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(m, 8, cv2.CV_32S)
            assert n == n_w and np.array_equal(lab, lab_w) and np.array_equal(st, st_w)
            assert np.array_equal(ce, ce_w, equal_nan=True)
    cv2.setNumThreads(nthreads)
    big = imgs.random_mask(700, 900, 3, 0.55)          # large enough for cv2's parallel labelling path
    n, lab, st, ce = oracle.ccl8_stats(big)
    n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(big, 8, cv2.CV_32S)
    assert n == n_w and np.array_equal(lab, lab_w) and np.array_equal(st, st_w) and np.array_equal(ce, ce_w)


def test_moments_variance():
    g = imgs.blurred_noise(211, 307, 1)
    s, ss, nz = oracle.moments_u8(g)
    assert s == int(g.sum(dtype=np.int64)) and nz == int(np.sum(g > 0))
    var = oracle.variance_from_moments(g.size, s, ss)
    assert abs(var - float(np.var(g))) < 1e-9 * var


def test_phash_close_to_float_dct():
    """The integer pHash tracks the usual float definition (scipy DCT of the 32x32 area mean)."""
    from scipy.fft import dct
    for seed in range(5):
        g = imgs.blurred_noise(320, 448, seed)
        h = oracle.phash(g)
        small = g.reshape(32, 10, 32, 14).mean((1, 3))
        d = dct(dct(small, axis=0, norm=None), axis=1, norm=None)[:8, :8]
        bits = (d > np.median(d)).reshape(-1)
        ref = 0
        for b in bits:
            ref = (ref << 1) | int(b)
        assert bin(h ^ ref).count("1") <= 2
    assert oracle.phash(imgs.blurred_noise(64, 64, 1)) != oracle.phash(imgs.blurred_noise(64, 64, 2))


def test_page_chain_matches_cv2():
    from synapta_image_segmentation_b200.synth import synth_page
    page, _ = synth_page(3, dpi=150, n_figures=2)
    r = cv2_chain.page_chain(page, 150)
    bs, c, k = cv2_chain.chain_params(150)
    assert (bs, c, k) == (25, 10, 21) and cv2_chain.chain_params(300) == (51, 10, 41)
    g = oracle.rgb2gray_cv(page)
    ink = oracle.adaptive_mean(g, bs, c, True) | oracle.canny(g, 50, 150)
    closed = oracle.morphology_ex(oracle.morph_rect(ink, 1, k, k), 3, k, k, 1)
    assert np.array_equal(closed, r["closed"])
    n, lab, st, ce = oracle.ccl8_stats(closed)
    assert n == r["n"] and np.array_equal(st, r["stats"]) and np.array_equal(ce, r["centroids"])


# ---- golden vectors generated from the imported reference ------------------------------------------
def _golden():
    p = os.path.join(GOLD, "reference_helpers.json")
    if not os.path.exists(p):
        pytest.skip("golden vectors not generated")
    return json.load(open(p))


def test_golden_reference_helpers_match_oracle_chain():
    """_detect_grid / _estimate_data_points fallback / variance / mask counts of the reference (run in the
    build container on the committed crops) equal the cv2-chain restatement and the C oracle."""
    gold = _golden()
    for name, rec in gold["crops"].items():
        img = np.array(Image.open(os.path.join(GOLD, name)))
        f = cv2_chain.crop_features(img)
        assert f["h_count"] == rec["h_count"] and f["v_count"] == rec["v_count"], name
        assert (f["h_count"] > 300 and f["v_count"] > 300) == rec["detect_grid"], name
        assert f["edge_px"] == rec["edge_px"]
        assert abs(f["variance"] - rec["variance"]) < 1e-9 * max(1.0, rec["variance"])
        assert f["mask_px"] == rec["mask_px"]
        g = cv2_chain.pil_gray(img)
        e = oracle.canny(g, 50, 150)
        assert int((e > 0).sum()) == rec["edge_px"]
        hl = oracle.morphology_ex(e, 2, 25, 1, 2); vl = oracle.morphology_ex(e, 2, 1, 25, 2)
        assert int((hl > 0).sum()) == rec["h_count"] and int((vl > 0).sum()) == rec["v_count"]
        s, ss, _ = oracle.moments_u8(g)
        assert abs(oracle.variance_from_moments(g.size, s, ss) - rec["variance"]) < 1e-9 * max(1.0, rec["variance"])
        if img.ndim == 3:
            assert oracle.hsv_mask(img)[1] == rec["mask_px"]


def test_colors_port_pinned_on_reference_outputs():
    """oracle/colors_port.py vs the imported reference's `_extract_dominant_colors` on the committed crops: the masked
    pixel count and the `[]` decision (pdf_image_segmentation.py:1574-1577) are identical; the colour lists themselves
    are a different (deterministic) clustering and are only required to be as long as the reference's."""
    from oracle import colors_port
    gold = _golden()
    for name, rec in gold["crops"].items():
        img = np.array(Image.open(os.path.join(GOLD, name)))
        n, cols, wts = colors_port.dominant_colors_hist(img)
        assert n == rec["mask_px"], name
        assert (len(cols) == 0) == (len(rec["dominant_colors_seed7"]) == 0), name
        assert len(cols) == len(rec["dominant_colors_seed7"]), name
        assert sum(wts) == (n if cols else 0), name
        if img.ndim == 3:
            _, hist, sums = colors_port.masked_histogram(img)
            n_c, hist_c, sums_c = oracle.hsv_hist(img, 4)          # the plain-C restatement of the histogram
            assert n_c == n and np.array_equal(hist_c.astype(np.int64), hist) and np.array_equal(sums_c.astype(np.int64), sums), name


def test_colors_port_known_answers():
    from oracle import colors_port
    flat = np.full((120, 200, 3), 255, np.uint8)
    flat[10:60, 10:90] = (200, 30, 40); flat[70:110, 20:180] = (30, 160, 60); flat[10:60, 100:190] = (40, 60, 210)
    assert colors_port.dominant_colors_hist(flat) == (14900, [(30, 160, 60), (40, 60, 210), (200, 30, 40)], [6400, 4500, 4000])
    assert colors_port.dominant_colors_hist(flat, n_colors=1) == (14900, [(78, 94, 99)], [14900])
    few = np.full((40, 40, 3), 128, np.uint8); few[0, :39] = (220, 20, 20); few[1:3, :30] = (220, 20, 20)
    assert colors_port.dominant_colors_hist(few) == (99, [], [])
    few[5, 0] = (20, 20, 220)
    assert colors_port.dominant_colors_hist(few) == (100, [(220, 20, 20), (20, 20, 220)], [99, 1])
    assert colors_port.dominant_colors_hist(np.zeros((30, 30), np.uint8)) == (0, [], [])
