"""Deterministic synthetic textbook pages -- shared by tests, bench and the CPU baseline.

White letter page, 1-inch margins, text-like dark word boxes on a 12 pt line pitch, and 0-3
"figures" per page (framed bar chart with optional grid, photo-like low-pass noise block, or a
box-and-arrow diagram).  About 10 % of the figures come from a small stock pool, so the same
figure re-appears on different pages at different positions: ground truth for the cross-page
perceptual-hash dedup.  Everything is a pure function of (base_seed, page_idx, dpi).

Figure sizes follow the reference's real crop distribution (investments_segmented/: median
699x457 px at 150 DPI) scaled by dpi/150.
"""
from __future__ import annotations

import numpy as np

PAGE_W_PT, PAGE_H_PT = 612.0, 792.0
STOCK_POOL = 16


def page_shape(dpi: int):
    """(H, W) of a letter page: 300 DPI -> (3300, 2550); 150 DPI -> (1650, 1275)."""
    return int(round(PAGE_H_PT * dpi / 72)), int(round(PAGE_W_PT * dpi / 72))


def _upsample(field: np.ndarray, h: int, w: int) -> np.ndarray:
    fy = np.linspace(0, field.shape[0] - 1, h)
    fx = np.linspace(0, field.shape[1] - 1, w)
    tmp = np.empty((field.shape[0], w))
    for i in range(field.shape[0]):
        tmp[i] = np.interp(fx, np.arange(field.shape[1]), field[i])
    y0 = np.floor(fy).astype(int)
    y1 = np.minimum(y0 + 1, field.shape[0] - 1)
    t = (fy - y0)[:, None]
    return tmp[y0] * (1 - t) + tmp[y1] * t


def render_figure(seed, dpi: int, max_h: int, max_w: int) -> np.ndarray:
    """One figure as an HxWx3 u8 array; pure function of (seed, dpi, max_h, max_w)."""
    rng = np.random.default_rng(seed)
    s = dpi / 150.0
    w = int(np.clip(rng.lognormal(np.log(699), 0.35), 240, 1150) * s)
    h = int(np.clip(rng.lognormal(np.log(457), 0.35), 180, 900) * s)
    w, h = min(w, max_w), min(h, max_h)
    kind = rng.choice(3, p=[0.6, 0.25, 0.15])
    img = np.full((h, w, 3), 255, np.uint8)
    ft = max(2, dpi // 75)                       # frame thickness: 4 px at 300 DPI
    if kind == 0:                                # framed bar chart
        img[:ft], img[-ft:], img[:, :ft], img[:, -ft:] = 0, 0, 0, 0
        if rng.random() < 0.6:                   # 25 pt grid of 1-px grey lines
            step = int(25 * dpi / 72)
            img[step::step, ft:-ft] = 180
            img[ft:-ft, step::step] = 180
        nb = int(rng.integers(6, 11))
        slot = (w - 2 * ft) / (nb + 1)
        for b in range(nb):
            bw = int(slot * 0.6)
            x0 = int(ft + slot * (b + 0.7))
            bh = int((h - 2 * ft) * rng.uniform(0.15, 0.9))
            img[h - ft - bh:h - ft, x0:x0 + bw] = rng.integers(0, 256, 3)
    elif kind == 1:                              # photo-like block, variance > 1500
        for c in range(3):
            low = rng.normal(0, 1, (max(4, h // 24), max(4, w // 24)))
            f = _upsample(low, h, w)
            f = (f - f.mean()) / (f.std() + 1e-9)
            img[:, :, c] = np.clip(128 + 60 * f, 0, 255).astype(np.uint8)
    else:                                        # boxes + connectors (diagram)
        nb = int(rng.integers(3, 7))
        centres = []
        bw, bh = w // 6, h // 6
        for _ in range(nb):
            cx = int(rng.integers(bw, w - bw)); cy = int(rng.integers(bh, h - bh))
            centres.append((cx, cy))
            x0, y0, x1, y1 = cx - bw // 2, cy - bh // 2, cx + bw // 2, cy + bh // 2
            img[y0:y1, x0:x1] = rng.integers(160, 256, 3)
            img[y0:y0 + ft, x0:x1] = 0; img[y1 - ft:y1, x0:x1] = 0
            img[y0:y1, x0:x0 + ft] = 0; img[y0:y1, x1 - ft:x1] = 0
        for a, b in zip(centres[:-1], centres[1:]):   # straight connectors, any angle
            n = max(abs(a[0] - b[0]), abs(a[1] - b[1])) + 1
            xs = np.linspace(a[0], b[0], n).astype(int); ys = np.linspace(a[1], b[1], n).astype(int)
            for d in range(ft):
                img[np.clip(ys + d, 0, h - 1), xs] = 40
    return img


def synth_page(page_idx: int, dpi: int = 300, base_seed: int = 1234, n_figures: int | None = None):
    """Returns (rgb u8 [H,W,3], truth) with truth = list of dict(box_px=(x0,y0,x1,y1), stock=int|None)."""
    h, w = page_shape(dpi)
    s = dpi / 72.0
    rng = np.random.default_rng([base_seed, page_idx])
    page = np.full((h, w, 3), 255, np.uint8)
    m = int(72 * s)                               # 1-inch margins
    if n_figures is None:
        n_figures = int(rng.choice(4, p=[0.58, 0.28, 0.10, 0.04]))   # mean ~0.6
    truth = []
    excl = []
    if n_figures:
        slot_h = (h - 2 * m) // n_figures
        pad = int(24 * s)
        for f in range(n_figures):
            stock = int(rng.integers(STOCK_POOL)) if rng.random() < 0.10 else None
            seed = [base_seed, 7_000_000 + stock] if stock is not None else [base_seed, page_idx, f]
            fig = render_figure(seed, dpi, slot_h - 2 * pad, w - 2 * m)
            fh, fw = fig.shape[:2]
            y0 = m + f * slot_h + pad + int(rng.integers(0, max(1, slot_h - 2 * pad - fh + 1)))
            x0 = m + int(rng.integers(0, max(1, w - 2 * m - fw + 1)))
            page[y0:y0 + fh, x0:x0 + fw] = fig
            truth.append(dict(box_px=(x0, y0, x0 + fw, y0 + fh), stock=stock))
            excl.append((x0 - pad, y0 - pad, x0 + fw + pad, y0 + fh + pad))
    pitch = int(np.ceil(12 * s)); wh = int(7 * s); gap = int(4 * s)
    for ry in range(m, h - m - wh, pitch):
        widths = (rng.uniform(5, 30, 64) * s).astype(int)
        shades = rng.integers(0, 61, 64)
        x = m
        for wd, sh in zip(widths, shades):
            if x + wd > w - m:
                break
            hit = False
            for (ex0, ey0, ex1, ey1) in excl:
                if x < ex1 and x + wd > ex0 and ry < ey1 and ry + wh > ey0:
                    hit = True
                    break
            if not hit:
                page[ry:ry + wh, x:x + wd] = sh
            x += wd + gap
    return page, truth


def dense_page(page_idx: int, dpi: int = 300, base_seed: int = 1234) -> np.ndarray:
    """A page without any blank paper: full-bleed low-pass "scan" background with per-pixel sensor noise, text boxes from edge
    to edge and two photo blocks.  It bounds the content-dependent cost of the stencils from above (their blank-row and
    constant-group shortcuts never fire); bench.py times it beside the text pages.  Pure function of (base_seed, page_idx, dpi)."""
    h, w = page_shape(dpi)
    s = dpi / 72.0
    rng = np.random.default_rng([base_seed, 900_000 + page_idx])
    low = rng.normal(0, 1, (max(4, h // 96), max(4, w // 96)))
    f = _upsample(low, h, w)
    f = (f - f.mean()) / (f.std() + 1e-9)
    page = np.empty((h, w, 3), np.uint8)
    for c in range(3):
        noise = rng.integers(-8, 9, (h, w))
        page[:, :, c] = np.clip(205 + 22 * f + noise, 0, 255).astype(np.uint8)
    pitch = int(np.ceil(12 * s)); wh = int(7 * s); gap = int(4 * s)
    for ry in range(gap, h - wh, pitch):
        widths = (rng.uniform(5, 30, 96) * s).astype(int)
        shades = rng.integers(0, 61, 96)
        x = gap
        for wd, sh in zip(widths, shades):
            if x + wd > w - gap:
                break
            page[ry:ry + wh, x:x + wd] = sh
            x += wd + gap
    for k in range(2):
        fig = render_figure([base_seed, 800_000 + page_idx, k], dpi, h // 3, w // 2)
        fh, fw = fig.shape[:2]
        y0 = int(rng.integers(0, h - fh)); x0 = int(rng.integers(0, w - fw))
        page[y0:y0 + fh, x0:x0 + fw] = fig
    return page


def synth_pages(n: int, dpi: int = 300, base_seed: int = 1234, start: int = 0, n_figures: int | None = None,
                out: np.ndarray | None = None) -> np.ndarray:
    """[n, H, W, 3] u8 batch (optionally into a caller-provided, e.g. pinned, array)."""
    h, w = page_shape(dpi)
    if out is None:
        out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        out[i] = synth_page(start + i, dpi, base_seed, n_figures)[0]
    return out
