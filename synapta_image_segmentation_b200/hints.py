"""Per-crop feature hints with the reference's signatures, computed on the GPU.

`FeatureHints` exposes the CV helpers of the reference's `OCRProcessor` under the same names, argument
meaning and return types (all @staticmethod, argument = a PIL image of any mode), so a maintainer can
monkey-patch them in (`OCRProcessor._detect_grid = FeatureHints._detect_grid`).

What runs where:
  GPU, bit-exact : grey (PIL / cv2 formula), Canny(50,150), OPEN(25x1|1x25, it=2) counts, chart-rule
                   OPEN counts, bar count via CCL stats, grey variance (integer moments), HSV mask count,
                   masked-pixel gather in numpy's `img[mask]` order, colour histogram.
  host, fed by the GPU's bit-exact edge map / grey / sample: cv2.HoughLinesP, cv2.findContours +
                   approxPolyDP, cv2.HoughCircles, cv2.SimpleBlobDetector, sklearn KMeans -- sequential
                   or randomised algorithms whose results are only reproducible by running the same code
                   on identical inputs.
"""
from __future__ import annotations

import re
from collections import defaultdict
from typing import Any, Dict, List, Optional

import numpy as np
import torch

try:                                   # Pillow >= 11.2 exports its pixel storage through the Arrow C data interface
    import pyarrow as _pa
except Exception:                      # pragma: no cover
    _pa = None

from .datamodel import ChartSpecificData, DiagramSpecificData, FigureSpecificData, ImageSpecificData, OCRResult
from .detector import get_context
from .ops import CLOSE, GRAY_CV, GRAY_PIL, OPEN, Context  # noqa: F401


def _to_device(image, ctx: Context):
    """PIL image -> (tensor, channels).  Mode 'L' stays grey (PIL's convert('L') is the identity there);
    everything else goes through RGB exactly like `image.convert('RGB')` / `convert('L')` would."""
    if image.mode == "L":
        a = np.array(image)
        return torch.from_numpy(a).to(ctx.device), 1
    if image.mode != "RGB":
        image = image.convert("RGB")
    a = np.array(image)
    return torch.from_numpy(a).to(ctx.device), 3


def _variance(ctx: Context, t: torch.Tensor, ch: int) -> float:
    m = ctx.moments(t, 1 if ch == 3 else 0).cpu().numpy()[0]
    n = t.shape[0] * t.shape[1]
    s1, s2 = int(m[0]), int(m[1])
    return (n * s2 - s1 * s1) / (n * n)


class FeatureHints:
    """GPU drop-ins for the CV helpers of OCRProcessor (pdf_image_segmentation.py:1320-1341, 1343-1461,
    1546-1617, 1695-1711, 1753-1810) and the old algorithm's feature drivers (old_algo:887-1010)."""

    # ---- shared front end: grey + Canny + line counts in one GPU pass ------------------------------
    @staticmethod
    def edge_features(image, want_edges: bool = False, kw: int = 25, kh: int = 25, gray_mode: int = GRAY_PIL) -> Dict[str, Any]:
        ctx = get_context()
        t, ch = _to_device(image, ctx)
        counts, edges = ctx.grid_counts(t, None, gray_mode, kw, kh, want_edges, channels=ch)
        c = counts.cpu().numpy()[0]
        out = dict(h_count=int(c[0]), v_count=int(c[1]), edge_px=int(c[2]))
        if want_edges:
            out["edges"] = np.ascontiguousarray(edges[0].cpu().numpy())
        return out

    @staticmethod
    def _detect_grid(image) -> bool:
        """pdf_image_segmentation.py:1546-1564"""
        f = FeatureHints.edge_features(image)
        return f["h_count"] > 300 and f["v_count"] > 300

    @staticmethod
    def _lines(edges: np.ndarray, min_len: int = 30, max_gap: int = 10):
        import cv2
        return cv2.HoughLinesP(edges, 1, np.pi / 180, threshold=50, minLineLength=min_len, maxLineGap=max_gap)

    @staticmethod
    def _count_arrows(image) -> int:
        """pdf_image_segmentation.py:1320-1341 (Hough on the GPU's Canny map)."""
        lines = FeatureHints._lines(FeatureHints.edge_features(image, want_edges=True)["edges"])
        if lines is None:
            return 0
        arrow_count = 0
        for line in lines:
            x1, y1, x2, y2 = line[0]
            angle = abs(np.arctan2(y2 - y1, x2 - x1) * 180 / np.pi)
            if 20 < angle < 70 or 110 < angle < 160:
                arrow_count += 1
        return min(arrow_count // 3, 20)

    @staticmethod
    def _extract_connections(image) -> List[Dict[str, Any]]:
        """pdf_image_segmentation.py:1695-1711"""
        lines = FeatureHints._lines(FeatureHints.edge_features(image, want_edges=True)["edges"])
        if lines is None:
            return []
        return [{"id": f"conn_{i}", "type": "arrow"} for i, _ in enumerate(lines[:20])]

    @staticmethod
    def _detect_shapes(image) -> Dict[str, int]:
        """pdf_image_segmentation.py:1753-1775; the `diamonds` branch is unreachable there too (always 0)."""
        import cv2
        shapes = {"rectangles": 0, "circles": 0, "diamonds": 0}
        edges = FeatureHints.edge_features(image, want_edges=True)["edges"]
        contours, _ = cv2.findContours(edges, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        for contour in contours:
            approx = cv2.approxPolyDP(contour, 0.04 * cv2.arcLength(contour, True), True)
            if len(approx) == 4:
                shapes["rectangles"] += 1
            elif len(approx) > 8:
                shapes["circles"] += 1
        return shapes

    @staticmethod
    def _detect_decision_points(image, ocr_result: Optional[OCRResult]) -> bool:
        """pdf_image_segmentation.py:1777-1789"""
        text = ocr_result.raw_text.lower() if ocr_result else ""
        has_keywords = any(kw in text for kw in ["if", "yes", "no", "decision", "choose", "select"])
        return has_keywords or FeatureHints._detect_shapes(image).get("diamonds", 0) > 0

    @staticmethod
    def _estimate_data_points(image) -> int:
        """pdf_image_segmentation.py:1596-1617: blob count if > 5, else min(edge_px // 150, 500)."""
        import cv2
        ctx = get_context()
        t, ch = _to_device(image, ctx)
        gray = t if ch == 1 else ctx.rgb2gray(t, GRAY_PIL)
        try:
            params = cv2.SimpleBlobDetector_Params()
            params.filterByArea = True
            params.minArea = 10
            params.maxArea = 150
            keypoints = cv2.SimpleBlobDetector_create(params).detect(np.ascontiguousarray(gray.cpu().numpy()))
            if len(keypoints) > 5:
                return len(keypoints)
        except Exception:
            pass
        counts, _ = ctx.grid_counts(gray, None, GRAY_PIL, 25, 25, False, channels=1)
        return min(int(counts[0, 2]) // 150, 500)

    @staticmethod
    def masked_pixel_sample(image, max_pixels: int = 5000) -> np.ndarray:
        """The pixel list the reference clusters (pdf_image_segmentation.py:1571-1583): RGB of the pixels with
        S>30 & V>40 & V<240 in raster order; more than `max_pixels` -> np.random.choice(n, max_pixels,
        replace=False) from numpy's GLOBAL generator, exactly like the reference (seed it to reproduce)."""
        ctx = get_context()
        if image.mode != "RGB":
            image = image.convert("RGB")
        t = torch.from_numpy(np.array(image)).to(ctx.device)
        res = ctx.hsv_mask_hist(t, None, want_hist=False, want_rows=True)
        n = int(res["count"][0])
        if n < 100:
            return np.zeros((0, 3), np.uint8)
        rows = res["row_count"][0].to(torch.int64)
        prefix = torch.cat([torch.zeros(1, dtype=torch.int64, device=ctx.device), torch.cumsum(rows, 0)])
        if n > max_pixels:
            idx = np.random.choice(n, max_pixels, replace=False)
        else:
            idx = np.arange(n)
        ranks = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(ctx.device)
        return ctx.hsv_mask_gather(t, None, prefix, ranks).cpu().numpy()

    @staticmethod
    def _extract_dominant_colors(image, n_colors: int = 5) -> List[str]:
        """pdf_image_segmentation.py:1566-1594: mask + sample on the GPU, KMeans(random_state=42, n_init=10) on host."""
        pixels = FeatureHints.masked_pixel_sample(image)
        if len(pixels) < 100:
            return []
        try:
            from sklearn.cluster import KMeans
            kmeans = KMeans(n_clusters=min(n_colors, len(pixels)), random_state=42, n_init=10)
            kmeans.fit(pixels)
            colors = kmeans.cluster_centers_.astype(int)
            return ["#%02x%02x%02x" % tuple(c) for c in colors]
        except ImportError:
            return []

    @staticmethod
    def decode_colors(row) -> Dict[str, Any]:
        """One row of `Context.colors_crops` -> dict(mask_px, dominant_colors ['#rrggbb'], color_weights [pixels])."""
        k = int(row[1])
        words = [int(v) for v in row[2:2 + k]]
        return dict(mask_px=int(row[0]), dominant_colors=["#%06x" % (v & 0xFFFFFF) for v in words],
                    color_weights=[v >> 24 for v in words])

    @staticmethod
    def dominant_colors_histogram(image, n_colors: int = 5, iters: int = 20) -> List[str]:
        """Deterministic GPU-histogram variant (an APPROXIMATION of the reference's KMeans over an unseeded random
        sample, pdf_image_segmentation.py:1581-1590, labelled as such): weighted Lloyd iterations over the exact
        centroids of the 4096 (R>>4,G>>4,B>>4) bins, started from the heaviest bins, all on the device
        (`synseg_colors_crops`).  The mask and the `fewer than 100 masked pixels -> []` decision (:1571-1577) are exact."""
        ctx = get_context()
        host, descs = FeatureHints.pack_crops([image])
        out, _ = ctx.colors_crops(host.to(ctx.device, non_blocking=True), descs, n_colors, iters)
        return FeatureHints.decode_colors(out[0].cpu().numpy())["dominant_colors"]

    @staticmethod
    def _detect_image_subtype(image, ocr_result: Optional[OCRResult]) -> Optional[str]:
        """pdf_image_segmentation.py:1791-1810"""
        text_length = len(ocr_result.raw_text) if ocr_result else 0
        if text_length > 500:
            return "scanned_page"
        elif text_length > 100:
            return "screenshot"
        ctx = get_context()
        t, ch = _to_device(image, ctx)
        return "photo" if _variance(ctx, t, ch) > 1500 else "illustration"

    @staticmethod
    def _detect_chart_subtype(image, ocr_result: Optional[OCRResult]) -> Optional[str]:
        """pdf_image_segmentation.py:1343-1461.  Text scoring as the reference; the visual signals come from the
        GPU (cv2-formula grey -> Canny -> OPEN counts -> CCL bar count), Hough transforms run on the host over the
        GPU's edge map / grey."""
        import cv2
        text = ocr_result.raw_text.lower() if ocr_result else ""
        ctx = get_context()
        if image.mode != "RGB":
            image = image.convert("RGB")
        t = torch.from_numpy(np.array(image)).to(ctx.device)
        height, width = t.shape[0], t.shape[1]
        gray = ctx.rgb2gray(t, GRAY_CV)
        scores = defaultdict(float)
        if re.search(r"\bpie\b", text) and "chart" in text:
            scores["pie"] += 3.0
        if "scatter" in text or "correlation" in text:
            scores["scatter"] += 3.0
        if "candlestick" in text or all(w in text for w in ["open", "close"]):
            scores["candlestick"] += 3.0
        if re.search(r"\bbar\b.*\bchart\b|\bbar\b.*\bgraph\b", text):
            scores["bar"] += 3.0
        if re.search(r"\bline\b.*\bchart\b|\bline\b.*\bgraph\b", text):
            scores["line"] += 3.0
        edges = ctx.canny(gray, 50, 150)
        v_detect = ctx.morph(edges, OPEN, 1, max(20, height // 20), iterations=2, binary=True)
        h_detect = ctx.morph(edges, OPEN, max(20, width // 20), 1, iterations=2, binary=True)
        mom = ctx.moments(torch.stack([v_detect.contiguous(), h_detect.contiguous()])).cpu().numpy()
        v_pixels, h_pixels = int(mom[0, 2]), int(mom[1, 2])
        edges_h = None
        if h_pixels > height * 8 and h_pixels > v_pixels * 1.5:
            scores["line"] += 2.5
            edges_h = np.ascontiguousarray(edges.cpu().numpy())
            lines = cv2.HoughLinesP(edges_h, 1, np.pi / 180, threshold=50, minLineLength=width // 4, maxLineGap=20)
            if lines is not None:
                long_h = sum(1 for line in lines if abs(line[0][3] - line[0][1]) < 10 and abs(line[0][2] - line[0][0]) > width * 0.2)
                if long_h >= 1:
                    scores["line"] += 1.5
        elif v_pixels > width * 10:
            scores["bar"] += 2.0
            # EXTERNAL contours + boundingRect height == 8-connected component bbox height
            n, _, stats, _ = ctx.ccl_stats(v_detect, max_labels=max(2, v_pixels + 2), want_labels=False)
            hts = stats[0, 1:int(n[0]), 3].cpu().numpy()
            if int((hts > height * 0.2).sum()) >= 3:
                scores["bar"] += 1.5
        if scores.get("line", 0) < 2.0 and scores.get("bar", 0) < 2.0:
            gray_h = np.ascontiguousarray(gray.cpu().numpy())
            m = min(width, height)
            circles = cv2.HoughCircles(gray_h, cv2.HOUGH_GRADIENT, dp=1, minDist=int(m * 0.3), param1=50, param2=50,
                                       minRadius=int(m * 0.2), maxRadius=int(m * 0.45))
            if circles is not None:
                large = [c for c in circles[0] if c[2] > m * 0.2]
                if len(large) == 1:
                    centre = large[0][:2].astype(int)
                    radius = int(large[0][2])
                    mask = np.zeros(gray_h.shape, dtype=np.uint8)
                    cv2.circle(mask, tuple(int(v) for v in centre), radius, 255, -1)
                    if edges_h is None:
                        edges_h = np.ascontiguousarray(edges.cpu().numpy())
                    density = np.sum(cv2.bitwise_and(edges_h, edges_h, mask=mask) > 0) / (np.pi * radius * radius)
                    if density > 0.015:
                        scores["pie"] += 2.5
        if scores:
            best = max(scores, key=scores.get)
            if scores[best] >= 2.0:
                return best
        return "unknown"          # pdf_image_segmentation.py:1461

    # ---- the old algorithm's per-type feature drivers (old_algo:887-1010), CV-derived fields only ----
    @staticmethod
    def process_chart_specific(image, ocr_result: Optional[OCRResult]) -> ChartSpecificData:
        d = ChartSpecificData()
        d.chart_subtype = FeatureHints._detect_chart_subtype(image, ocr_result)
        d.series_count = len(d.legend_items) if d.legend_items else 1
        d.grid_detected = FeatureHints._detect_grid(image)
        d.color_scheme = FeatureHints._extract_dominant_colors(image)
        d.estimated_data_points = FeatureHints._estimate_data_points(image)
        return d

    @staticmethod
    def process_diagram_specific(image, ocr_result: Optional[OCRResult]) -> DiagramSpecificData:
        d = DiagramSpecificData()
        d.connections = FeatureHints._extract_connections(image)
        d.arrow_count = ocr_result.detected_arrows if ocr_result else 0
        d.shapes_detected = FeatureHints._detect_shapes(image)
        d.has_decision_points = FeatureHints._detect_decision_points(image, ocr_result)
        return d

    @staticmethod
    def process_image_specific(image, ocr_result: Optional[OCRResult]) -> ImageSpecificData:
        d = ImageSpecificData()
        d.image_subtype = FeatureHints._detect_image_subtype(image, ocr_result)
        if ocr_result and ocr_result.raw_text:
            d.contains_text = len(ocr_result.raw_text.strip()) > 10
            n = len(ocr_result.raw_text)
            d.text_density = "dense" if n > 500 else ("moderate" if n > 100 else ("sparse" if n > 0 else "none"))
        d.dominant_colors = FeatureHints._extract_dominant_colors(image)
        return d

    @staticmethod
    def process_figure_specific(image, ocr_result: Optional[OCRResult]) -> FigureSpecificData:
        d = FigureSpecificData()
        if ocr_result and ocr_result.raw_text:
            matches = re.findall(r"\([a-z]\)|\b[a-z]\)", ocr_result.raw_text.lower())
            if len(matches) >= 2:
                d.is_composite = True
                d.sub_figure_count = len(matches)
        d.contains_chart = FeatureHints._detect_grid(image)
        d.contains_diagram = (ocr_result.detected_arrows if ocr_result else 0) > 3
        ctx = get_context()
        t, ch = _to_device(image, ctx)
        d.contains_image = _variance(ctx, t, ch) > 1000
        return d

    # ---- batched form for many crops (config 4) -----------------------------------------------------
    @staticmethod
    def crop_view(image):
        """PIL image or uint8 array -> (2-D uint8 view [height, width * channels], width, height, channels, keep-alive).
        PIL 'RGB' and 'L' images are handed over WITHOUT conversion through Pillow's Arrow export (zero copy): RGB
        storage is 4 bytes per pixel (RGBX), which `synseg_hints_crops` reads directly (channels = 4).  Images Pillow
        cannot export that way (multi-block storage of very large images, other modes) go through np.asarray."""
        if isinstance(image, np.ndarray):
            if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] not in (3, 4)):
                raise ValueError("crops must be uint8 [H, W], [H, W, 3] or [H, W, 4] arrays or PIL images")
            ch = 1 if image.ndim == 2 else image.shape[2]
            h, w = image.shape[0], image.shape[1]
            return image.reshape(h, w * ch), w, h, ch, image
        if image.mode not in ("RGB", "L"):
            image = image.convert("RGB")
        w, h = image.size
        ch = 1 if image.mode == "L" else 4
        image.load()                           # lazily opened files; memory-mapped / buffer-backed images stay read-only
        if _pa is not None and hasattr(image, "__arrow_c_array__") and not image.readonly:   # (Pillow 12.2 crashes exporting those)
            try:
                arr = _pa.array(image)
                flat = (arr.flatten() if ch == 4 else arr).to_numpy(zero_copy_only=True)
                if flat.size == h * w * ch:
                    return flat.reshape(h, w * ch), w, h, ch, (arr, image)
            except Exception:
                pass
        a = np.asarray(image)
        ch = 1 if a.ndim == 2 else 3
        return a.reshape(h, w * ch), w, h, ch, a

    @staticmethod
    def pack_crops(crops):
        """Crops -> (pinned uint8 buffer, [(offset, width, height, row_stride, channels)]): the packed layout
        `synseg_hints_crops` reads, rows padded to 16 bytes.  One buffer for the whole list (see `hints_batch` for the
        streamed form)."""
        views = [FeatureHints.crop_view(c) for c in crops]
        descs, off = [], 0
        for v, w, h, ch, _ in views:
            rs = (w * ch + 15) // 16 * 16
            descs.append((off, w, h, rs, ch))
            off += rs * h
        if not descs:
            return None, []
        host = torch.empty(off, dtype=torch.uint8, pin_memory=True)
        _copy_views(host.numpy(), views, descs)
        return host, descs

    @staticmethod
    def hints_batch(crops) -> List[Dict[str, Any]]:
        """One dict per crop with the deterministic GPU hint quantities: h_count, v_count, edge_px,
        grid_detected, variance, mask_px, data_points_fallback, image_subtype_visual, dominant_colors
        (deterministic histogram clustering, see `dominant_colors_histogram`) and color_weights.

        Streamed: the crops are cut into chunks of <= 256 MB; a chunk is copied (thread pool, no pixel conversion for
        PIL images) into one of two persistent pinned slots, sent to the device and processed by one
        `synseg_hints_crops` call while the host fills the other slot; results come back once at the end."""
        ctx = get_context()
        views = [FeatureHints.crop_view(c) for c in crops]
        if not views:
            return []
        stager = _stager(ctx)
        chunks, cur, off = [], [], 0
        for v, w, h, ch, _ in views:
            rs = (w * ch + 15) // 16 * 16
            if cur and off + rs * h > stager.slot_bytes:
                chunks.append((cur, off))
                cur, off = [], 0
            cur.append((off, w, h, rs, ch))
            off += rs * h
        chunks.append((cur, off))
        results, colours, first = [], [], 0
        for k, (descs, nbytes) in enumerate(chunks):
            host, dev, ev = stager.slot(k, nbytes)
            if ev is not None:
                ev.synchronize()                       # the copy that last read this pinned slot has finished
            _copy_views(host.numpy(), views[first:first + len(descs)], descs, stager.pool)
            dev[:nbytes].copy_(host[:nbytes], non_blocking=True)
            stager.mark(k)
            results.append(ctx.hints_crops(dev[:nbytes], descs))
            colours.append(ctx.colors_crops(dev[:nbytes], descs)[0])
            first += len(descs)
        res = torch.cat(results).cpu().numpy()
        col = torch.cat(colours).cpu().numpy()
        return [FeatureHints._hint_dict(r, w * h, cr) for (v, w, h, ch, _), r, cr in zip(views, res, col)]


    @staticmethod
    def hints_regions(pages: torch.Tensor, regions_per_page, with_colors: bool = True) -> List[List[Dict[str, Any]]]:
        """`hints_batch` for regions of pages that are still on the device (the normal case right after detection): per page,
        per region dict (as `RasterRegionDetector.detect_regions_batch` returns them: 'crop_px' = x, y, w, h) the same hint
        dict as `hints_batch` gives for the cropped PIL image -- read in place by `synseg_hints_rois` / `synseg_colors_rois`,
        no crop is cut, packed or uploaded.  pages: CUDA u8 [B,H,W,3] (or grey [B,H,W]: no colours)."""
        ctx = get_context(pages.device.index)
        rois = [(i,) + tuple(r["crop_px"]) for i, regs in enumerate(regions_per_page) for r in regs]
        if not rois:
            return [[] for _ in regions_per_page]
        grey = pages.dim() == 3 and not (pages.shape[-1] == 3 and pages.stride(-2) == 3)
        res = ctx.hints_rois(pages, rois).cpu().numpy()
        col = ctx.colors_rois(pages, rois).cpu().numpy() if (with_colors and not grey) else None
        out, j = [], 0
        for regs in regions_per_page:
            cur = []
            for r in regs:
                _, _, w, h = r["crop_px"]
                cur.append(FeatureHints._hint_dict(res[j], w * h, col[j] if col is not None else None))
                j += 1
            out.append(cur)
        return out

    @staticmethod
    def _hint_dict(r, n: int, cr) -> Dict[str, Any]:
        s1, s2 = int(r[3]), int(r[4])
        var = (n * s2 - s1 * s1) / (n * n)
        d = dict(h_count=int(r[0]), v_count=int(r[1]), edge_px=int(r[2]), grid_detected=bool(r[0] > 300 and r[1] > 300),
                 variance=var, mask_px=int(r[6]), data_points_fallback=min(int(r[2]) // 150, 500),
                 image_subtype_visual="photo" if var > 1500 else "illustration")
        if cr is not None:
            c = FeatureHints.decode_colors(cr)
            d.update(dominant_colors=c["dominant_colors"], color_weights=c["color_weights"])
        return d


def _copy_views(hv: np.ndarray, views, descs, pool=None) -> None:
    """Copies every crop view into its place of the packed buffer (numpy releases the GIL for these copies)."""
    def one(i):
        v = views[i][0]
        o, w, h, rs, ch = descs[i]
        hv[o:o + rs * h].reshape(h, rs)[:, :w * ch] = v
    if pool is None or len(descs) < 4:
        for i in range(len(descs)):
            one(i)
    else:
        list(pool.map(one, range(len(descs)), chunksize=max(1, len(descs) // (4 * pool._max_workers))))


class _CropStager:
    """Two persistent pinned slots + device buffers for `hints_batch` (no allocation in the steady state)."""

    def __init__(self, ctx: Context, slot_bytes: int = 256 << 20):
        import concurrent.futures as cf
        import os
        self.device = ctx.device
        self.slot_bytes = slot_bytes
        self.host = [None, None]
        self.dev = [None, None]
        self.events = [None, None]
        self.pool = cf.ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))

    def slot(self, k: int, nbytes: int):
        s = k & 1
        if self.host[s] is None or self.host[s].numel() < nbytes:
            n = max(nbytes, self.slot_bytes)
            if self.events[s] is not None:
                self.events[s].synchronize()
            self.host[s] = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            self.dev[s] = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self.host[s], self.dev[s], self.events[s]

    def mark(self, k: int) -> None:
        ev = torch.cuda.Event()
        ev.record()
        self.events[k & 1] = ev


_STAGERS: Dict[int, "_CropStager"] = {}


def _stager(ctx: Context) -> _CropStager:
    st = _STAGERS.get(id(ctx))
    if st is None:
        st = _STAGERS[id(ctx)] = _CropStager(ctx)
    return st
