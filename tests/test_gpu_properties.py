"""Property tests (hypothesis) of the CUDA primitives and the page pipeline against the live cv2 primitives: random shapes
(1 x 1 upwards, widths around the 16 / 32 / 64 / 480-pixel unit boundaries of the kernels), random content classes (blank,
noise, sparse specks, solid blocks, lines), random parameters.  Bar: bit-exact.  SURVEY.md section 4 item (2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")
hyp = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large],
              derandomize=True)

widths = st.one_of(st.integers(1, 70), st.sampled_from([95, 96, 97, 127, 128, 129, 255, 256, 257, 479, 480, 481, 511, 513, 960, 961]))
heights = st.one_of(st.integers(1, 70), st.sampled_from([127, 128, 129, 200]))


@st.composite
def grey_images(draw, binary=False):
    h, w = draw(heights), draw(widths)
    kind = draw(st.sampled_from(["noise", "blank", "specks", "blocks", "lines", "smooth"]))
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 31)))
    if kind == "noise":
        a = rng.integers(0, 256, (h, w))
    elif kind == "blank":
        a = np.full((h, w), int(rng.integers(0, 256)))
    elif kind == "specks":
        a = np.full((h, w), 255)
        a[rng.random((h, w)) < 0.02] = 0
    elif kind == "blocks":
        a = np.full((h, w), 255)
        for _ in range(int(rng.integers(1, 6))):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            a[y:y + int(rng.integers(1, max(2, h // 2))), x:x + int(rng.integers(1, max(2, w // 2)))] = int(rng.integers(0, 256))
    elif kind == "lines":
        a = np.full((h, w), 255)
        a[::max(1, int(rng.integers(2, 9))), :] = int(rng.integers(0, 200))
        a[:, ::max(1, int(rng.integers(2, 9)))] = int(rng.integers(0, 200))
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        a = (127 + 120 * np.sin(yy / rng.uniform(2, 15)) * np.cos(xx / rng.uniform(2, 15))).astype(int)
    a = a.astype(np.uint8)
    if binary:
        a = ((a < 128) * 255).astype(np.uint8)
    return a


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@settings(max_examples=60, **COMMON)
@given(img=grey_images(), bs=st.sampled_from([3, 5, 11, 25, 51, 101]), c=st.integers(-20, 40), inv=st.booleans())
def test_adaptive_threshold_property(ctx, img, bs, c, inv):
    want = cv2.adaptiveThreshold(img, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV if inv else cv2.THRESH_BINARY, bs, c)
    assert np.array_equal(ctx.adaptive_mean(dev(img), bs, c, inv).cpu().numpy(), want)


@settings(max_examples=60, **COMMON)
@given(img=grey_images(), lo=st.integers(0, 120), span=st.integers(0, 200))
def test_canny_property(ctx, img, lo, span):
    want = cv2.Canny(img, lo, lo + span)
    assert np.array_equal(ctx.canny(dev(img), lo, lo + span).cpu().numpy(), want)


@settings(max_examples=80, **COMMON)
@given(img=grey_images(binary=True), op=st.sampled_from([0, 1, 2, 3]), kw=st.integers(1, 45), kh=st.integers(1, 45), it=st.integers(1, 3),
       centre=st.booleans(), data=st.data())
def test_morphology_property(ctx, img, op, kw, kh, it, centre, data):
    ax, ay = (-1, -1) if centre else (data.draw(st.integers(0, kw - 1)), data.draw(st.integers(0, kh - 1)))
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
    cvop = [cv2.MORPH_ERODE, cv2.MORPH_DILATE, cv2.MORPH_OPEN, cv2.MORPH_CLOSE][op]
    want = cv2.morphologyEx(img, cvop, se, anchor=(ax, ay), iterations=it)
    for binary in (True, False):
        got = ctx.morph(dev(img), op, kw, kh, (ax, ay), it, binary=binary).cpu().numpy()
        assert np.array_equal(got, want), (op, kw, kh, ax, ay, it, binary)


@settings(max_examples=60, **COMMON)
@given(img=grey_images(binary=True))
def test_connected_components_property(ctx, img):
    n_w, lab_w, st_w, ce_w = cv2.connectedComponentsWithStats(img, 8, cv2.CV_32S)
    n, lab, stats, cent = ctx.ccl_stats(dev(img), max_labels=max(8, n_w + 2))
    assert int(n[0]) == n_w
    assert np.array_equal(lab[0].cpu().numpy(), lab_w)
    assert np.array_equal(stats[0, :n_w].cpu().numpy(), st_w)
    if img.any() and not img.all():
        assert np.array_equal(cent[0, :n_w].cpu().numpy(), ce_w)


@st.composite
def rgb_pages(draw):
    h = draw(st.sampled_from([33, 64, 97, 150, 211]))
    w = draw(st.sampled_from([33, 160, 333, 480, 497, 523, 700, 977]))        # 3 w bytes per row: every alignment of the TMA boxes
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 31)))
    page = np.full((h, w, 3), 255, np.uint8)
    for _ in range(int(rng.integers(1, 8))):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        hh, ww = int(rng.integers(1, max(2, h // 2))), int(rng.integers(1, max(2, w // 2)))
        page[y:y + hh, x:x + ww] = rng.integers(0, 256, 3) if rng.random() < 0.7 else rng.integers(0, 256, page[y:y + hh, x:x + ww].shape)
    if rng.random() < 0.3:
        page = rng.integers(0, 256, page.shape).astype(np.uint8)
    return np.ascontiguousarray(page)


@settings(max_examples=40, **COMMON)
@given(page=rgb_pages(), bs=st.sampled_from([5, 13, 25]), k=st.integers(1, 14), batch=st.integers(1, 3))
def test_page_pipeline_property(ctx, page, bs, k, batch):
    """The fused pipeline (TMA-fed front end included) == the cv2 chain, for any page size and row alignment."""
    g = cv2.cvtColor(page, cv2.COLOR_RGB2GRAY)
    thr = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, bs, 10)
    ink = cv2.bitwise_or(thr, cv2.Canny(g, 50, 150))
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
    closed = cv2.morphologyEx(cv2.dilate(ink, se), cv2.MORPH_CLOSE, se)
    n_w, _, st_w, ce_w = cv2.connectedComponentsWithStats(closed, 8, cv2.CV_32S)
    pages = dev(np.stack([page] * batch))
    n, stats, cent = ctx.detect_pages(pages, bs, 10, k, max_labels=max(16, n_w + 2))
    for j in range(batch):
        assert int(n[j]) == n_w
        assert np.array_equal(stats[j, :n_w].cpu().numpy(), st_w)
        if closed.any() and not closed.all():
            assert np.array_equal(cent[j, :n_w].cpu().numpy(), ce_w)


def _weak_chain_image(h, w, seed, vertical):
    """Long WEAK edges (step of 20 grey levels: |gradient| = 80, between 50 and 150) that cross many 32-row bands / 32-pixel
    words and touch a STRONG edge (step of 120) at one end only, plus weak edges that touch nothing strong (must vanish)."""
    rng = np.random.default_rng(seed)
    a = np.full((h, w), 200, np.uint8)
    if vertical:
        for x in range(20, w - 20, 37):
            a[10:h - 10, x:x + 9] = 180                      # weak stripes over the whole height
        a[10:14, :] = 60                                     # strong bar across the top: connected stripes survive
        a[h // 2:h // 2 + 3, 20:w - 20:74] = 200             # cut every second stripe: its lower half hangs on nothing strong
    else:
        for y in range(20, h - 20, 23):
            a[y:y + 7, 10:w - 10] = 180
        a[:, 10:14] = 60
        a[20:h - 20:46, w // 2:w // 2 + 3] = 200
    return a


@pytest.mark.parametrize("vertical", [True, False])
@pytest.mark.parametrize("hw", [(700, 500), (333, 1300)])
def test_hysteresis_sweeps_and_union_find_fallback(ctx, hw, vertical, monkeypatch):
    """Weak chains longer than the propagation sweeps reach (hyst_sweep.cu) fall through to the union-find; both paths and the
    sweeps-off path give cv2.Canny's edges."""
    h, w = hw
    img = _weak_chain_image(h, w, 1, vertical)
    want = cv2.Canny(img, 50, 150)
    assert 0 < int((want > 0).sum())
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("SYNSEG_NO_HYST_SWEEPS", env)
        else:
            monkeypatch.delenv("SYNSEG_NO_HYST_SWEEPS", raising=False)
        counts, edges = ctx.grid_counts(dev(img), None, 1, 25, 25, True, channels=1)
        assert np.array_equal(edges[0].cpu().numpy(), want), env
        assert int(counts[0, 2]) == int((want > 0).sum())
    monkeypatch.delenv("SYNSEG_NO_HYST_SWEEPS", raising=False)
    # through the page pipeline (edges OR-ed into the threshold plane): same components with and without the sweeps
    rgb = dev(np.repeat(img[None, :, :, None], 3, axis=3))
    a = ctx.detect_pages(rgb, 15, 10, 5, max_labels=4096)
    monkeypatch.setenv("SYNSEG_NO_HYST_SWEEPS", "1")
    b = ctx.detect_pages(rgb, 15, 10, 5, max_labels=4096)
    monkeypatch.delenv("SYNSEG_NO_HYST_SWEEPS", raising=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1][0, :int(a[0][0])], b[1][0, :int(a[0][0])])


@settings(max_examples=60, **COMMON)
@given(img=grey_images(), lo=st.integers(0, 120), span=st.integers(0, 200))
def test_canny_bit_path_property(ctx, img, lo, span):
    """The bit-plane Canny path (propagation sweeps + union-find behind them) that the pipelines use, against cv2.Canny(50, 150)."""
    want = cv2.Canny(img, 50, 150)
    _, edges = ctx.grid_counts(dev(img), None, 1, 25, 25, True, channels=1)
    assert np.array_equal(edges[0].cpu().numpy(), want)
