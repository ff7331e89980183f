"""GPU parity: every CUDA primitive, called through the C ABI, against the CPU oracle
(oracle/synseg_oracle.c) and the live cv2 / PIL primitive on the same seeded inputs.
Bar: bit-exact (integer / byte / index work)."""
import numpy as np
import pytest

import imgs

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")
import oracle  # noqa: E402
from oracle import cv2_chain  # noqa: E402

SIZES = [(1, 1), (2, 3), (7, 5), (31, 33), (64, 64), (97, 131), (240, 317), (513, 770)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------------ grey
@pytest.mark.parametrize("mode", [0, 1])
def test_gray_exhaustive_256cubed(ctx, mode):
    """All 16.7M RGB triples, both fixed-point formulas."""
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    rgb = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    got = host(ctx.rgb2gray(dev(rgb), mode))
    want = oracle.rgb2gray_cv(rgb) if mode == 0 else oracle.rgb2gray_pil(rgb)
    assert np.array_equal(got, want)
    if mode == 0:
        assert np.array_equal(got, cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    else:
        from PIL import Image
        assert np.array_equal(got, np.array(Image.fromarray(rgb, "RGB").convert("L")))


@pytest.mark.parametrize("hw", SIZES + [(3300 // 4, 2550)])
def test_gray_shapes_and_alignment(ctx, hw):
    """Tightly packed rows of any width (row base alignment 0..15), batch of 2, both modes."""
    h, w = hw
    rgb = np.stack([imgs.rgb_noise(h, w, 1), imgs.rgb_noise(h, w, 2)])
    for mode, f in ((0, oracle.rgb2gray_cv), (1, oracle.rgb2gray_pil)):
        got = host(ctx.rgb2gray(dev(rgb), mode))
        assert np.array_equal(got, f(rgb))
    # unaligned base pointer: slice one byte column off a larger buffer
    if w > 2:
        d = dev(rgb)[:, :, 1:, :]
        assert np.array_equal(host(ctx.rgb2gray(d, 0)), oracle.rgb2gray_cv(rgb[:, :, 1:, :]))
        out = torch.empty((2, h, w + 5), dtype=torch.uint8, device="cuda")[:, :, 3:3 + w - 1]
        ctx.rgb2gray(d, 0, out=out)
        assert np.array_equal(host(out), oracle.rgb2gray_cv(rgb[:, :, 1:, :]))


# ------------------------------------------------------------------------------------------------ adaptive
@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("bs,c,inv", [(3, 2, True), (15, 10, True), (25, 10, True), (51, 10, True), (31, 5, False), (101, 7, True)])
def test_adaptive_mean(ctx, hw, bs, c, inv):
    h, w = hw
    for k, g in enumerate((imgs.blurred_noise(h, w, 5 + k) if k else imgs.shapes(max(h, 8), max(w, 8), 3)[:h, :w] for k in range(2))):
        g = np.ascontiguousarray(g)
        got = host(ctx.adaptive_mean(dev(g), bs, c, inv))
        want = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV if inv else cv2.THRESH_BINARY, bs, c)
        assert np.array_equal(got, want), (hw, bs, c, inv, int((got != want).sum()))


def test_adaptive_mean_batch_and_oracle(ctx):
    g = np.stack([imgs.blurred_noise(300, 411, s) for s in range(3)])
    got = host(ctx.adaptive_mean(dev(g), 25, 10, True))
    for i in range(3):
        assert np.array_equal(got[i], oracle.adaptive_mean(g[i], 25, 10, True))


def test_adaptive_mean_wide_strips(ctx):
    """Width > 4096 - halos exercises the multi-strip path."""
    g = imgs.blurred_noise(70, 5000, 11, passes=1)
    got = host(ctx.adaptive_mean(dev(g), 51, 10, True))
    assert np.array_equal(got, cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 51, 10))


# ------------------------------------------------------------------------------------------------ canny
@pytest.mark.parametrize("hw", SIZES)
def test_canny(ctx, hw):
    h, w = hw
    for g in (imgs.blurred_noise(h, w, 21, passes=2), imgs.shapes(max(h, 8), max(w, 8), 22)[:h, :w]):
        g = np.ascontiguousarray(g)
        got = host(ctx.canny(dev(g), 50, 150))
        want = cv2.Canny(g, 50, 150)
        assert np.array_equal(got, want), (hw, int((got != want).sum()))
        assert np.array_equal(got, oracle.canny(g, 50, 150))


def test_canny_long_weak_chain(ctx):
    """A long weak ramp edge seeded by one strong pixel: hysteresis must propagate along the whole chain."""
    h, w = 64, 1500
    g = np.full((h, w), 100, np.uint8)
    g[32:, :] = 120            # weak horizontal edge (Sobel magnitude 80: between 50 and 150)
    g[32:, :6] = 200           # strong start
    got = host(ctx.canny(dev(g), 50, 150))
    want = cv2.Canny(g, 50, 150)
    assert want[:, 1000:].any()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("passes", [1, 2, 3, 4, 6])
def test_canny_weak_strong_mixtures(ctx, passes):
    """Blur strength sweeps the weak / strong ratio of the kept pixels (hysteresis with skipped strong-strong unions)."""
    for seed, (h, w) in enumerate([(300, 401), (257, 640), (129, 2100)]):
        g = np.ascontiguousarray(imgs.blurred_noise(h, w, 40 + seed, passes=passes))
        for lo, hi in ((50, 150), (20, 60), (5, 200)):
            assert np.array_equal(host(ctx.canny(dev(g), lo, hi)), cv2.Canny(g, lo, hi)), (passes, h, w, lo, hi)


def test_canny_rows_full_of_candidates(ctx):
    """Strip rows with hundreds of candidates take the packed vertical-class pass (|dy| >= 3 |dx|) before the candidate list:
    horizontal edges of every contrast with slanted, noisy and tapering stretches (classes mixed inside one row), plateaus that tie
    with the row above / below (m > up but m >= down), edges on the first and last image rows, widths of several strips."""
    rng = np.random.default_rng(77)
    for h, w in ((97, 481), (160, 1700), (64, 2550)):
        g = np.full((h, w), 230, np.uint8)
        y = 3
        while y + 6 < h:
            c = int(rng.integers(0, 200))
            t = int(rng.integers(1, 5))
            g[y:y + t, :] = c                                                    # a bar: two horizontal edges, equal magnitudes on two rows
            if rng.random() < 0.5:                                               # a slanted stretch: the direction class changes along the row
                x0 = int(rng.integers(0, w - 40))
                for i in range(40):
                    g[y + (i // 8) % (t + 1):y + t, x0 + i] = c
            if rng.random() < 0.5:                                               # noise on top of a stretch
                x0 = int(rng.integers(0, w - 64))
                g[y - 1:y + t + 1, x0:x0 + 64] = rng.integers(0, 256, (t + 2, 64))
            if rng.random() < 0.5:                                               # a horizontal ramp along the bar: |dx| > 0 everywhere
                g[y:y + t, :] = np.clip(c + (np.arange(w) * int(rng.integers(1, 4))) % 120, 0, 255).astype(np.uint8)
            y += t + int(rng.integers(2, 7))
        g[0, :] = 10; g[h - 1, : w // 2] = 20                                    # edges against the zero ring outside the image
        for lo, hi in ((50, 150), (10, 30), (100, 400)):
            got = host(ctx.canny(dev(g), lo, hi))
            assert np.array_equal(got, cv2.Canny(g, lo, hi)), (h, w, lo, hi)


def test_canny_batch(ctx):
    g = np.stack([imgs.shapes(333, 517, s) for s in range(4)])
    got = host(ctx.canny(dev(g)))
    for i in range(4):
        assert np.array_equal(got[i], cv2.Canny(g[i], 50, 150))


# ------------------------------------------------------------------------------------------------ morphology
KERNELS = [(3, 3), (1, 25), (25, 1), (21, 21), (41, 41), (20, 1), (1, 30), (4, 6), (81, 3)]


@pytest.mark.parametrize("hw", [(1, 1), (5, 7), (40, 33), (97, 131), (240, 317)])
@pytest.mark.parametrize("kw,kh", KERNELS)
def test_morph_binary_ops(ctx, hw, kw, kh):
    h, w = hw
    m = imgs.random_mask(h, w, 31, 0.7)
    m2 = cv2.Canny(imgs.shapes(max(h, 8), max(w, 8), 5)[:h, :w].copy(), 50, 150)
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
    for src in (m, m2):
        d = dev(src)
        assert np.array_equal(host(ctx.morph(d, 1, kw, kh, binary=True)), cv2.dilate(src, se))
        assert np.array_equal(host(ctx.morph(d, 0, kw, kh, binary=True)), cv2.erode(src, se))
        for op, cvop in ((2, cv2.MORPH_OPEN), (3, cv2.MORPH_CLOSE)):
            for it in (1, 2):
                got = host(ctx.morph(d, op, kw, kh, iterations=it, binary=True))
                assert np.array_equal(got, cv2.morphologyEx(src, cvop, se, iterations=it)), (hw, kw, kh, op, it)


@pytest.mark.parametrize("kw,kh", [(251, 1), (1, 401), (300, 400), (227, 385), (226, 384), (99, 191)])
def test_morph_binary_huge_kernels(ctx, kw, kh):
    """Kernel sizes around and beyond the limits of the register row pass (k <= 226) and the van Herk column pass
    (k <= 384): the shared-memory fallback kernels and the 12-word / 128-thread variants."""
    m = imgs.random_mask(450, 613, 77, 0.97)
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
    d = dev(m)
    assert np.array_equal(host(ctx.morph(d, 1, kw, kh, binary=True)), cv2.dilate(m, se))
    assert np.array_equal(host(ctx.morph(d, 0, kw, kh, binary=True)), cv2.erode(m, se))
    inv = 255 - m
    assert np.array_equal(host(ctx.morph(dev(inv), 3, kw, kh, binary=True)), cv2.morphologyEx(inv, cv2.MORPH_CLOSE, se))


@pytest.mark.parametrize("hw", [(1, 1), (5, 7), (40, 33), (97, 131), (200, 300)])
@pytest.mark.parametrize("kw,kh", [(3, 3), (1, 25), (25, 1), (21, 21), (20, 1), (4, 6), (49, 1)])
def test_morph_grey(ctx, hw, kw, kh):
    h, w = hw
    src = imgs.blurred_noise(h, w, 41, passes=1)
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
    d = dev(src)
    assert np.array_equal(host(ctx.morph(d, 1, kw, kh)), cv2.dilate(src, se))
    assert np.array_equal(host(ctx.morph(d, 0, kw, kh)), cv2.erode(src, se))
    assert np.array_equal(host(ctx.morph(d, 2, kw, kh, iterations=2)), cv2.morphologyEx(src, cv2.MORPH_OPEN, se, iterations=2))
    assert np.array_equal(host(ctx.morph(d, 3, kw, kh)), cv2.morphologyEx(src, cv2.MORPH_CLOSE, se))


def test_morph_anchor_and_oracle(ctx):
    src = imgs.random_mask(60, 80, 77, 0.3)
    for (kw, kh, ax, ay) in [(5, 1, 0, 0), (5, 1, 4, 0), (6, 4, 1, 3), (1, 9, 0, 8)]:
        for op in (0, 1):
            want = oracle.morph_rect(src, op, kw, kh, ax, ay)
            se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
            cvw = (cv2.dilate if op else cv2.erode)(src, se, anchor=(ax, ay))
            assert np.array_equal(want, cvw)
            assert np.array_equal(host(ctx.morph(dev(src), op, kw, kh, anchor=(ax, ay), binary=True)), want)
            assert np.array_equal(host(ctx.morph(dev(src), op, kw, kh, anchor=(ax, ay), binary=False)), want)


# ------------------------------------------------------------------------------------------------ CCL
def _ccl_check(ctx, m, max_labels=None):
    n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(m, 8, cv2.CV_32S)
    cap = max_labels or max(n_w, 1)
    n, lab, st, ce = ctx.ccl_stats(dev(m), cap)
    assert int(n[0]) == n_w
    assert np.array_equal(host(lab)[0], lab_w)
    assert np.array_equal(host(st)[0, :n_w], st_w)
    got_c = host(ce)[0, :n_w]
    assert np.array_equal(np.isnan(got_c), np.isnan(ce_w))
    assert np.array_equal(got_c[~np.isnan(got_c)], ce_w[~np.isnan(ce_w)])


@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("density", [0.05, 0.3, 0.5, 0.62, 0.9])
def test_ccl_random(ctx, hw, density):
    _ccl_check(ctx, imgs.random_mask(*hw, seed=int(density * 100), density=density))


def test_ccl_adversarial(ctx):
    _ccl_check(ctx, imgs.spiral_mask(201, 333))
    _ccl_check(ctx, imgs.checkerboard(64, 65))
    _ccl_check(ctx, np.zeros((17, 19), np.uint8))
    _ccl_check(ctx, np.full((17, 19), 255, np.uint8))
    m = np.zeros((6, 14), np.uint8)
    for (r, c) in [(1, 0), (0, 10), (4, 3), (3, 12)]:     # SURVEY.md Appendix A label-order vector
        m[r, c] = 255
    _ccl_check(ctx, m)
    rows = np.zeros((101, 257), np.uint8); rows[::2] = 255   # many 1-px rows
    _ccl_check(ctx, rows)
    cols = np.zeros((101, 257), np.uint8); cols[:, ::2] = 1  # any non-zero value is foreground
    _ccl_check(ctx, cols)


def test_ccl_batch_no_labels_and_capacity(ctx):
    ms = np.stack([imgs.random_mask(120, 150, s, 0.4) for s in range(3)])
    n, lab, st, ce = ctx.ccl_stats(dev(ms), 4096, want_labels=False)
    assert lab is None
    for i in range(3):
        n_w, _, st_w, ce_w = cv2.connectedComponentsWithStats(ms[i], 8, cv2.CV_32S)
        assert int(n[i]) == n_w
        assert np.array_equal(host(st)[i, :n_w], st_w)
        assert np.array_equal(host(ce)[i, :n_w], ce_w)
    n2, _, st2, _ = ctx.ccl_stats(dev(ms[0]), 8)
    n_w, _, st_w, _ = cv2.connectedComponentsWithStats(ms[0], 8, cv2.CV_32S)
    assert int(n2[0]) == -n_w               # capacity exceeded is reported, not silently truncated
    assert np.array_equal(host(st2)[0, :8], st_w[:8])


# ------------------------------------------------------------------------------------------------ reductions
def test_moments_and_variance(ctx):
    g = np.stack([imgs.blurred_noise(211, 307, s) for s in range(2)])
    got = host(ctx.moments(dev(g)))
    for i in range(2):
        s, ss, nz = oracle.moments_u8(g[i])
        assert tuple(int(v) for v in got[i]) == (s, ss, nz)
        var = oracle.variance_from_moments(g[i].size, s, ss)
        assert abs(var - float(np.var(g[i]))) <= 1e-9 * max(1.0, var)
    rgb = imgs.rgb_noise(90, 140, 3)
    rois = [(0, 0, 0, 140, 90), (0, 13, 7, 50, 31), (0, 139, 89, 1, 1)]
    for kind, f in ((1, oracle.rgb2gray_pil), (2, oracle.rgb2gray_cv)):
        got = host(ctx.moments(dev(rgb), kind, rois))
        gray = f(rgb)
        for (_, x, y, w, h), row in zip(rois, got):
            assert tuple(int(v) for v in row) == oracle.moments_u8(np.ascontiguousarray(gray[y:y + h, x:x + w]))


def test_hsv_mask_exhaustive_and_hist(ctx):
    r, g, b = np.meshgrid(np.arange(0, 256, 1, dtype=np.uint8), np.arange(0, 256, 3, dtype=np.uint8), np.arange(0, 256, 1, dtype=np.uint8), indexing="ij")
    rgb = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(256 * 86, 256, 3))
    res = ctx.hsv_mask_hist(dev(rgb), want_hist=True, want_sums=True, want_rows=True)
    want_mask = cv2_chain.hsv_mask(rgb)
    assert int(res["count"][0]) == int(want_mask.sum())
    assert np.array_equal(host(res["row_count"])[0], want_mask.sum(1))
    n, hist, sums = oracle.hsv_hist(rgb, 4)
    assert n == int(want_mask.sum())
    assert np.array_equal(host(res["hist"])[0].astype(np.uint32), hist)
    assert np.array_equal(host(res["chan_sum"])[0].astype(np.uint64), sums)


def test_hsv_gather_matches_numpy_mask_order(ctx):
    rgb = imgs.rgb_noise(120, 200, 9)
    roi = (0, 10, 5, 150, 100)
    crop = rgb[5:105, 10:160]
    mask = cv2_chain.hsv_mask(crop)
    pixels = crop[mask].reshape(-1, 3)
    res = ctx.hsv_mask_hist(dev(rgb), [roi], want_hist=False, want_rows=True)
    rows = host(res["row_count"])[0].astype(np.int64)
    assert int(res["count"][0]) == len(pixels)
    prefix = np.concatenate([[0], np.cumsum(rows)]).astype(np.uint64)
    ranks = np.random.default_rng(0).choice(len(pixels), 500, replace=False).astype(np.int64)
    got = host(ctx.hsv_mask_gather(dev(rgb), roi, dev(prefix.view(np.int64)), dev(ranks)))
    assert np.array_equal(got, pixels[ranks])


# ------------------------------------------------------------------------------------------------ pHash
def test_phash_matches_oracle_and_dedup(ctx):
    crops = [imgs.blurred_noise(h, w, s) for s, (h, w) in enumerate([(64, 64), (457, 699), (33, 1200), (20, 17), (1, 1), (300, 31)])]
    hashes = []
    for c in crops:
        got = int(host(ctx.phash(dev(c)))[0]) & 0xFFFFFFFFFFFFFFFF
        assert got == oracle.phash(c)
        hashes.append(got)
    # RGB region through PIL grey == grey crop hash
    rgb = imgs.rgb_noise(200, 300, 4)
    roi = (0, 20, 30, 199, 150)
    gray = oracle.rgb2gray_pil(rgb)[30:180, 20:219]
    got = int(host(ctx.phash(dev(rgb), 1, [roi]))[0]) & 0xFFFFFFFFFFFFFFFF
    assert got == oracle.phash(np.ascontiguousarray(gray))
    # dedup: duplicates of entry 1 and near-duplicate (2 bits flipped) of entry 0
    h = np.array([hashes[0], hashes[1], hashes[1], hashes[0] ^ 0b101, hashes[2]], dtype=np.uint64)
    keys = np.array([5, 1, 9, 7, 3], dtype=np.uint64)
    keep = host(ctx.phash_dedup(dev(h.view(np.int64)), dev(keys.view(np.int64)), 4))
    assert keep.tolist() == [1, 1, 0, 0, 1]


def test_strided_and_offset_planes(ctx):
    """Views into larger buffers: row stride != width (aligned and unaligned strides), base pointers off by 1..3 bytes,
    pitched outputs -- for the stencils with vector paths -- and the grey plane handed back by the fused pipeline."""
    h, w = 150, 203
    big = imgs.blurred_noise(h + 4, 256, 9, passes=2)
    for x0 in (0, 1, 3, 16):
        g = big[2:2 + h, x0:x0 + w]
        t = dev(big)[2:2 + h, x0:x0 + w]                       # row stride 256 (16-byte aligned), base offset x0
        assert not t.is_contiguous()
        assert np.array_equal(host(ctx.canny(t, 50, 150)), cv2.Canny(np.ascontiguousarray(g), 50, 150)), x0
        got = host(ctx.adaptive_mean(t, 25, 10, True))
        want = cv2.adaptiveThreshold(np.ascontiguousarray(g), 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 25, 10)
        assert np.array_equal(got, want), x0
        m = (g > 128).astype(np.uint8) * 255
        tm = dev(np.where(big > 128, 255, 0).astype(np.uint8))[2:2 + h, x0:x0 + w]
        n, lab, st, ce = ctx.ccl_stats(tm, 4096)
        n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(np.ascontiguousarray(m), 8, cv2.CV_32S)
        assert int(n[0]) == n_w and np.array_equal(host(lab[0]), lab_w) and np.array_equal(host(st[0, :n_w]), st_w)
    # fused pipeline with the grey plane handed back (cv2 grey, kept for downstream features)
    rgb = imgs.rgb_noise(97, 131, 3)
    gray_out = torch.empty((1, 97, 144), dtype=torch.uint8, device="cuda")[:, :, :131]
    n, st, ce = ctx.detect_pages(dev(rgb)[None], 15, 5, 5, max_labels=4096, gray_out=gray_out)
    assert np.array_equal(host(gray_out[0]), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    n2, st2, _ = ctx.detect_pages(dev(rgb)[None], 15, 5, 5, max_labels=4096)
    assert int(n[0]) == int(n2[0]) and torch.equal(st[0, :int(n[0])], st2[0, :int(n[0])])
