"""The canonical page chain through the live cv2/PIL wheels -- TEST INFRASTRUCTURE ONLY.

This is "the reference's CPU OpenCV path" of BASELINE.json: the cv2 4.13.0 primitives the GPU
kernels must match bit for bit, composed as SURVEY.md 8(d) states:

    gray (COLOR_RGB2GRAY) -> adaptiveThreshold(MEAN_C, BINARY_INV, bs, C) -> Canny(50,150)
    -> ink = thr | edges -> dilate(rect k) -> morphologyEx(CLOSE, rect k)
    -> connectedComponentsWithStats(8)

and the reference's per-crop helper arithmetic (S:1320-1341, S:1546-1617, S:1753-1810) restated
through the same wheels so it can run on the GPU box where /root/reference is absent.
Used by tests (as the checker) and by bench.py's cpu_baseline / --impl reference legs (as the
thing timed on the host cores).  Never imported by the product.
"""
from __future__ import annotations

import time

import cv2
import numpy as np
from PIL import Image


def chain_params(dpi: int):
    """(bs, C, k) of SURVEY.md 8a B2/B4: bs=51,k=41 @300 DPI; bs=25,k=21 @150 DPI."""
    return (dpi // 6) | 1, 10, int(10 * dpi / 72) | 1


def page_chain(rgb: np.ndarray, dpi: int, timings: dict | None = None):
    """Returns dict(gray, thr, edges, ink, dil, closed, n, labels, stats, centroids)."""
    bs, c, k = chain_params(dpi)
    t = [time.perf_counter()]

    def lap(name):
        t.append(time.perf_counter())
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + (t[-1] - t[-2])

    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY); lap("gray")
    thr = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, bs, c); lap("adaptive")
    edges = cv2.Canny(gray, 50, 150); lap("canny")
    ink = cv2.bitwise_or(thr, edges); lap("or")
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
    dil = cv2.dilate(ink, se); lap("dilate")
    closed = cv2.morphologyEx(dil, cv2.MORPH_CLOSE, se); lap("close")
    n, labels, stats, cent = cv2.connectedComponentsWithStats(closed, 8, cv2.CV_32S); lap("ccl")
    return dict(gray=gray, thr=thr, edges=edges, ink=ink, dil=dil, closed=closed,
                n=n, labels=labels, stats=stats, centroids=cent)


def pil_gray(rgb_or_gray: np.ndarray) -> np.ndarray:
    """np.array(Image.convert('L')) -- the grey the reference's helpers use (S:1323 et al.)."""
    if rgb_or_gray.ndim == 2:
        return rgb_or_gray
    return np.array(Image.fromarray(rgb_or_gray, "RGB").convert("L"))


def grid_counts(gray_pil: np.ndarray):
    """(h_count, v_count, edges) of _detect_grid, S:1546-1564."""
    edges = cv2.Canny(gray_pil, 50, 150)
    h_kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (25, 1))
    v_kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1, 25))
    h_lines = cv2.morphologyEx(edges, cv2.MORPH_OPEN, h_kernel, iterations=2)
    v_lines = cv2.morphologyEx(edges, cv2.MORPH_OPEN, v_kernel, iterations=2)
    return int(np.sum(h_lines > 0)), int(np.sum(v_lines > 0)), edges


def chart_counts(rgb: np.ndarray):
    """(v_pixels, h_pixels, vertical_bars) of _detect_chart_subtype's visual part, S:1365-1376,1403-1404."""
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    height, width = gray.shape
    edges = cv2.Canny(gray, 50, 150)
    v_kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1, max(20, height // 20)))
    v_detect = cv2.morphologyEx(edges, cv2.MORPH_OPEN, v_kernel, iterations=2)
    h_kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (max(20, width // 20), 1))
    h_detect = cv2.morphologyEx(edges, cv2.MORPH_OPEN, h_kernel, iterations=2)
    contours, _ = cv2.findContours(v_detect, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    bars = sum(1 for c in contours if cv2.boundingRect(c)[3] > height * 0.2)
    return int(np.sum(v_detect > 0)), int(np.sum(h_detect > 0)), int(bars)


def hsv_mask(rgb: np.ndarray) -> np.ndarray:
    """Boolean mask of _extract_dominant_colors, S:1571-1575."""
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    return (hsv[:, :, 1] > 30) & (hsv[:, :, 2] > 40) & (hsv[:, :, 2] < 240)


def crop_features(rgb_or_gray: np.ndarray) -> dict:
    """The deterministic per-crop quantities of SURVEY.md 8a C1,C3,C4,C8 through cv2/PIL/numpy."""
    g = pil_gray(rgb_or_gray)
    h_count, v_count, edges = grid_counts(g)
    out = dict(h_count=h_count, v_count=v_count, edge_px=int(np.sum(edges > 0)),
               variance=float(np.var(g)))
    if rgb_or_gray.ndim == 3:
        out["mask_px"] = int(hsv_mask(rgb_or_gray).sum())
    else:
        out["mask_px"] = 0
    return out
