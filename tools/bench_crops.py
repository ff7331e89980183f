"""Throughput of the batched crop-hint path (BASELINE.json configs[3]): N synthetic crops with the size distribution
of the reference's shipped crops (median 699x457 at 150 DPI, up to 1191x1500) through `synseg_hints_crops`.

Phases, each timed separately:  pack (PIL -> pinned, host), H2D (pinned -> device), resident (the C-ABI call on the
packed device buffer, CUDA events), and the whole `FeatureHints.hints_batch` call from PIL images.  The crop-by-crop
path (SYNSEG_HINTS_PER_CROP=1) and the cv2 / PIL / numpy chain on one host core are timed beside it on a sample.

    python tools/bench_crops.py [n_crops] [--json out.json]
"""
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cv2_chain  # noqa: E402
from synapta_image_segmentation_b200.detector import get_context  # noqa: E402
from synapta_image_segmentation_b200.hints import FeatureHints  # noqa: E402
from synapta_image_segmentation_b200.synth import render_figure  # noqa: E402


def crop_sizes(n, seed=4):
    """(h, w) pairs: log-normal around the reference's median crop 457 x 699, clipped to [60, 1500] x [70, 1191]."""
    rng = np.random.default_rng(seed)
    h = np.clip(np.exp(rng.normal(np.log(457), 0.45, n)), 60, 1500).astype(int)
    w = np.clip(np.exp(rng.normal(np.log(699), 0.35, n)), 70, 1191).astype(int)
    return list(zip(h.tolist(), w.tolist()))


def timed_events(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 2000
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    sizes = crop_sizes(n)
    n_uniq = min(n, 48)
    uniq = [render_figure([11, i], 150, *sizes[i]) for i in range(n_uniq)]            # distinct contents, cycled
    crops = []
    for i, (h, w) in enumerate(sizes):
        u = uniq[i % n_uniq]
        if u.shape[:2] != (h, w):                                                    # tile / cut the figure to this crop's size
            u = np.tile(u, (-(-h // u.shape[0]), -(-w // u.shape[1]), 1))[:h, :w]
        crops.append(Image.fromarray(np.ascontiguousarray(u)))
    ctx = get_context()
    FeatureHints.hints_batch(crops[:32])                                             # warm-up (arena, pinned pool)
    torch.cuda.synchronize()

    t0 = time.perf_counter()
    host, descs = FeatureHints.pack_crops(crops)
    t_pack = time.perf_counter() - t0
    dev = torch.empty_like(host, device=ctx.device)
    ms_h2d = timed_events(lambda: dev.copy_(host, non_blocking=True))
    ctx.hints_crops(dev, descs)                                                      # sizes the arena
    launches0 = ctx.launches
    ms_res = timed_events(lambda: ctx.hints_crops(dev, descs))
    launches = (ctx.launches - launches0) // 3
    ctx.profile_begin()                                                              # per-kernel CUDA events inside the library
    res = ctx.hints_crops(dev, descs).cpu().numpy()
    per_kernel = {}
    for name, ms in ctx.profile_end():
        per_kernel[name] = per_kernel.get(name, 0.0) + ms
    per_kernel = {k: round(v, 3) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}
    # dominant colours of the same packed batch (synseg_colors_crops: one CTA per crop, histogram + clustering on chip)
    ctx.colors_crops(dev, descs)
    ms_col = timed_events(lambda: ctx.colors_crops(dev, descs))
    col = ctx.colors_crops(dev, descs)[0].cpu().numpy()
    m = min(n, 200)
    host3, descs3 = FeatureHints.pack_crops([np.asarray(c) for c in crops[:m]])      # the crop-by-crop path reads RGB only
    dev3 = host3.to(ctx.device)
    os.environ["SYNSEG_HINTS_PER_CROP"] = "1"
    ctx.hints_crops(dev3, descs3)
    ms_per_crop = timed_events(lambda: ctx.hints_crops(dev3, descs3), reps=2) / m
    res_pc = ctx.hints_crops(dev3, descs3).cpu().numpy()
    del os.environ["SYNSEG_HINTS_PER_CROP"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    FeatureHints.hints_batch(crops)
    torch.cuda.synchronize()
    t_api = time.perf_counter() - t0

    k = min(n, 48)
    t0 = time.perf_counter()
    ref = [cv2_chain.crop_features(np.asarray(crops[i])) for i in range(k)]
    t_cpu = (time.perf_counter() - t0) / k
    ok = all(int(res[i][0]) == ref[i]["h_count"] and int(res[i][1]) == ref[i]["v_count"] and int(res[i][2]) == ref[i]["edge_px"]
             and int(res[i][6]) == ref[i]["mask_px"] for i in range(k))
    same = bool(np.array_equal(res[:m], res_pc))
    from oracle import colors_port
    kc = min(n, 12)
    t0 = time.perf_counter()
    ref_c = [colors_port.dominant_colors_hist(np.asarray(crops[i])) for i in range(kc)]
    t_cpu_col = (time.perf_counter() - t0) / kc
    ok_col = all(FeatureHints.decode_colors(col[i])["dominant_colors"] == ["#%02x%02x%02x" % c for c in ref_c[i][1]]
                 and int(col[i][0]) == ref_c[i][0] for i in range(kc))
    mpx = sum(h * w for h, w in sizes) / 1e6
    gb = host.numel() / 1e9
    out = dict(n_crops=n, megapixels=round(mpx, 1), packed_gb=round(gb, 3), layout="RGBX" if descs[0][4] == 4 else "RGB",
               resident_ms=round(ms_res, 3), resident_crops_per_s=round(n / ms_res * 1e3, 1), resident_gpx_per_s=round(mpx / ms_res, 2),
               kernel_launches=int(launches), per_crop_path_ms_per_crop=round(ms_per_crop, 4),
               per_crop_path_crops_per_s=round(1e3 / ms_per_crop, 1),
               h2d_ms=round(ms_h2d, 3), h2d_gb_per_s=round(gb / ms_h2d * 1e3, 2),
               one_buffer_pack_s_incl_pinned_alloc=round(t_pack, 3), api_s=round(t_api, 3), api_crops_per_s=round(n / t_api, 1),
               cv2_chain_one_core_ms_per_crop=round(t_cpu * 1e3, 3), cv2_chain_one_core_crops_per_s=round(1 / t_cpu, 1),
               colors_resident_ms=round(ms_col, 3), colors_crops_per_s=round(n / ms_col * 1e3, 1),
               colors_gb_per_s=round(gb / ms_col * 1e3, 1), colors_port_one_core_ms_per_crop=round(t_cpu_col * 1e3, 3),
               colors_parity_vs_port=ok_col,
               parity_vs_cv2_chain=ok, ragged_equals_per_crop_path=same, resident_ms_per_kernel=per_kernel)
    print(json.dumps(out))
    if out_json:
        with open(out_json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
