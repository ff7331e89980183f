"""CPU restatement of the deterministic dominant-colour clustering -- TEST INFRASTRUCTURE ONLY.

The reference (pdf_image_segmentation.py:1566-1594, `_extract_dominant_colors`) masks the crop with
cv2's HSV (S > 30 & V > 40 & V < 240, :1571-1574), returns [] below 100 masked pixels (:1577) and
otherwise runs sklearn KMeans over an UNSEEDED random sample of the masked pixels (:1581-1590), so its
colour list is not reproducible.  The product (`csrc/colors.cu`, `synseg_colors_crops`) reproduces
the mask, the count and the `[]` decision exactly and replaces the sampled KMeans by a deterministic
clustering of the exact 4096-bin colour histogram.  This file states that clustering in numpy int64
(fixed point, 1/64 of a grey level; no floating point anywhere), operation for operation, so the GPU
result can be checked bit for bit; the mask comes from the live cv2 wheel (`cv2_chain.hsv_mask`).
Never imported by the product.
"""
from __future__ import annotations

import numpy as np

from . import cv2_chain

BINS = 4096


def masked_histogram(rgb: np.ndarray):
    """(masked pixel count, hist int64[4096], channel sums int64[4096, 3]) of an RGB uint8 image."""
    mask = cv2_chain.hsv_mask(rgb)
    px = rgb[mask].reshape(-1, 3).astype(np.int64)
    bins = ((px[:, 0] >> 4) << 8) | ((px[:, 1] >> 4) << 4) | (px[:, 2] >> 4)
    hist = np.bincount(bins, minlength=BINS).astype(np.int64)
    sums = np.zeros((BINS, 3), np.int64)
    for c in range(3):
        np.add.at(sums[:, c], bins, px[:, c])
    return int(mask.sum()), hist, sums


Q = 64      # fixed point of the clustering: 1/64 of a grey level


def dominant_colors_hist(rgb: np.ndarray, n_colors: int = 5, iters: int = 20, min_pixels: int = 100):
    """(mask_px, [(r, g, b)], [cluster pixel counts]) -- the integer arithmetic spelled out in csrc/colors.cu."""
    if rgb.ndim == 2:                              # grey crop: convert('RGB') has S = 0 everywhere
        return 0, [], []
    n, hist, sums = masked_histogram(rgb)
    if n < min_pixels:
        return n, [], []
    nz = np.nonzero(hist)[0]
    w, cs = hist[nz], sums[nz]
    pts = (Q * cs + (w // 2)[:, None]) // w[:, None]          # bin centroids, rounded to 1/64 (int64)
    k = min(n_colors, len(nz))
    start = np.lexsort((nz, -w))[:k]               # heaviest bins first, ties to the lower bin index
    centres = pts[start].copy()
    cnt = np.zeros(k, np.int64)
    for _ in range(iters):
        diff = pts[:, None, :] - centres[None, :, :]
        d = (diff * diff).sum(-1)                  # exact integers (< 2^31)
        a = d.argmin(1)                            # first minimum = lower centre index on ties
        cnt = np.zeros(k, np.int64)
        tot = np.zeros((k, 3), np.int64)
        np.add.at(cnt, a, w)
        np.add.at(tot, a, w[:, None] * pts)
        new = centres.copy()
        for j in range(k):
            if cnt[j]:
                new[j] = (tot[j] + cnt[j] // 2) // cnt[j]
        moved = not np.array_equal(new, centres)
        centres = new
        if not moved:
            break
    cols = centres // Q                            # truncation, like the reference's `.astype(int)` (:1591)
    return n, [tuple(int(v) for v in c) for c in cols], [int(v) for v in cnt]
