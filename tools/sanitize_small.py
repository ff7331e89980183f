"""Small workload for `compute-sanitizer --tool memcheck` (run on the GPU box): every kernel family once, odd sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synapta_image_segmentation_b200.ops import Context  # noqa: E402
from synapta_image_segmentation_b200.synth import synth_page  # noqa: E402

ctx = Context(0)
rng = np.random.default_rng(0)
for (h, w) in ((97, 131), (240, 317), (33, 2081)):
    rgb = torch.from_numpy(rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)).cuda()
    g = ctx.rgb2gray(rgb, 0)
    ctx.rgb2gray(rgb, 1)
    ctx.adaptive_mean(g, 15, 5, True)
    ctx.adaptive_mean(g, 51, 10, False)
    e = ctx.canny(g, 50, 150)
    for (kw, kh) in ((3, 3), (41, 41), (1, 30), (81, 3)):
        ctx.morph(e, 3, kw, kh, binary=True)
        ctx.morph(e, 2, kw, kh, iterations=2, binary=True)
    ctx.morph(g, 1, 5, 7)
    ctx.ccl_stats(e, 4096, want_labels=True)
    ctx.moments(g)
    ctx.hsv_mask_hist(rgb, want_sums=True, want_rows=True)
    ctx.phash(rgb, 1)
    ctx.detect_pages(rgb, 15, 5, 5, max_labels=4096)
page, _ = synth_page(0, 72, n_figures=2)
t = torch.from_numpy(page).cuda()[None]
n, stats, _ = ctx.detect_pages(t, 13, 10, 11, max_labels=256)
packed = torch.from_numpy(page.reshape(-1)).cuda()
ctx.hints_crops(packed, [(0, 612, 300, 612 * 3, 3), (612 * 3 * 300, 200, 100, 612 * 3, 3)])
torch.cuda.synchronize()
print("sanitize workload done", int(n[0]))
ctx.close()
