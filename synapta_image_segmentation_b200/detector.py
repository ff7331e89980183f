"""Raster candidate-region detector: the drop-in for the inside of the reference's region detection.

The reference finds regions from the PDF object model (captions, drawing commands, embedded images:
pdf_image_segmentation.py:2763-2849, 3105-3146, 3511-3594).  This stage takes already-rasterised pages
instead (the `_render_region` handoff format, :3638-3657: RGB u8, HWC, px = pt * dpi / 72) and emits the
same region-dict / VisualSegment records:

  GPU (libsynseg.so): grey -> adaptive threshold | Canny -> dilate -> close -> 8-connected components
                      with bbox / area / centroid per component   (bit-exact with the cv2 chain)
  host (this file)  : component boxes -> points -> the reference's own filter / cluster / merge rules
                      (geometry.py) -> _validate_embedded_image scoring with GPU grey variance.

Large components become regions directly ("raster_cc", the analogue of an embedded-image rect); the
small leftovers are clustered with the drawing-command rule (<100 pt from the seed, >=3 members, pad 10 pt,
5000 < area < 0.8 page) into "raster_cluster" regions, merged with `_detect_visual_regions`' duplicate rule.
"""
from __future__ import annotations

import hashlib
import io
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import geometry as G
from .datamodel import BoundingBox, VisualSegment, VisualType
from .ops import Context


@dataclass
class DetectConfig:
    """Chain parameters derived from the raster DPI unless given."""
    dpi: int = 300
    block_size: Optional[int] = None     # (dpi // 6) | 1        -> 51 @300, 25 @150
    C: int = 10
    k: Optional[int] = None              # int(10 * dpi / 72) | 1 -> 41 @300, 21 @150
    canny_lo: int = 50                   # pdf_image_segmentation.py:1324
    canny_hi: int = 150
    max_labels: int = 1024
    min_extent_pt: float = 50.0          # :3450 (drawing / image regions must exceed 50 x 50 pt)
    keep_score: float = G.KEEP_SCORE     # :2885

    def resolved(self):
        bs = self.block_size or ((self.dpi // 6) | 1)
        k = self.k or (int(10 * self.dpi / 72) | 1)
        return bs, self.C, k


_default_ctx: Optional[Context] = None


def get_context(device: Optional[int] = None) -> Context:
    """Process-wide context for the current CUDA device (one context per (process, GPU))."""
    global _default_ctx
    if _default_ctx is None or (device is not None and _default_ctx.device.index != device):
        _default_ctx = Context(device)
    return _default_ctx


class RasterRegionDetector:
    def __init__(self, config: Optional[DetectConfig] = None, ctx: Optional[Context] = None, device: Optional[int] = None):
        self.cfg = config or DetectConfig()
        self.ctx = ctx or get_context(device)

    # ---- GPU stage ---------------------------------------------------------------------------
    def detect_components(self, pages: torch.Tensor, out=None):
        """pages: CUDA u8 [B,H,W,3] (or [H,W,3]).  Returns device tensors (n_labels, stats, centroids)."""
        bs, c, k = self.cfg.resolved()
        return self.ctx.detect_pages(pages, bs, c, k, self.cfg.canny_lo, self.cfg.canny_hi, self.cfg.max_labels, out=out)

    # ---- host stage --------------------------------------------------------------------------
    def candidate_regions(self, stats: np.ndarray, n_labels: int, page_width_pt: float, page_height_pt: float,
                          priors: Optional[Sequence[Dict]] = None) -> List[Dict]:
        """Component stats of one page ([n,5] = x,y,w,h,area in px; row 0 = background) -> region dicts
        (bbox in points) before validation.

        priors: regions a PDF object model produced for this page (the reference's caption-based regions,
        `_detect_by_captions`, pdf_image_segmentation.py:3148-3254: dicts with 'bbox' and, when known, 'caption_bbox').
        They take the place `caption_regions` has in `_detect_visual_regions` (:3114-3144): every prior is kept, a raster
        region is added unless more than half of it lies inside a prior or it sits right above a prior's caption
        (SURVEY.md 8f rank 4)."""
        if n_labels < 0:
            raise RuntimeError(f"page has {-n_labels} components, more than max_labels={self.cfg.max_labels}")
        s = 72.0 / self.cfg.dpi
        page_area = page_width_pt * page_height_pt
        primary: List[Dict] = []
        small: List[List[float]] = []
        for (x, y, w, h, area) in stats[1:n_labels].tolist():
            rect = [x * s, y * s, (x + w) * s, (y + h) * s]
            bbox = BoundingBox(rect[0], rect[1], rect[2], rect[3], page_width_pt, page_height_pt)
            a = bbox.area()
            big = (G.DRAWING_MIN_AREA < a < page_area * G.DRAWING_MAX_PAGE_FRACTION
                   and (rect[2] - rect[0]) > self.cfg.min_extent_pt and (rect[3] - rect[1]) > self.cfg.min_extent_pt)
            if big:
                primary.append({"bbox": bbox, "caption": None, "detection_method": "raster_cc",
                                "notes": f"Connected component of {area} px"})
            elif a < page_area * G.DRAWING_MAX_PAGE_FRACTION:
                small.append(rect)
        secondary = G.regions_from_rects(small, page_width_pt, page_height_pt, "raster_cluster", "raster components")
        raster = G.merge_visual_regions(primary, secondary)
        if not priors:
            return raster
        # every raster region is tested against the PRIORS only (among themselves the raster regions were merged above:
        # with no priors the result must not change), by the two rules of _detect_visual_regions
        priors = list(priors)
        return priors + [r for r in raster
                         if not G.overlaps_with_existing(r["bbox"], priors)
                         and not any("caption_bbox" in p and G.caption_near_region(r["bbox"], p["caption_bbox"]) for p in priors)]

    def _crop_px(self, bbox: BoundingBox, width: int, height: int):
        x, y, w, h = bbox.to_pixels(self.cfg.dpi)
        x = min(max(x, 0), width - 1); y = min(max(y, 0), height - 1)
        w = max(1, min(w, width - x)); h = max(1, min(h, height - y))
        return x, y, w, h

    def detect_regions_batch(self, pages: torch.Tensor, page_nums: Optional[Sequence[int]] = None,
                             page_width_pt: Optional[float] = None, page_height_pt: Optional[float] = None,
                             with_hash: bool = False, priors: Optional[Sequence[Optional[Sequence[Dict]]]] = None) -> List[List[Dict]]:
        """RGB pages (CUDA u8 [B,H,W,3]) -> per page a list of region dicts, sorted by (y0, x0).

        Region dict = the reference's schema (pdf_image_segmentation.py:3246-3252 / 3550-3555):
        {'bbox': BoundingBox (points), 'caption': None, 'detection_method', 'notes'} plus
        'confidence' / 'validation' (the _validate_embedded_image score and notes), 'page_num',
        'crop_px' (x, y, w, h) and, with_hash, 'phash'."""
        if pages.dim() == 3:
            pages = pages.unsqueeze(0)
        b, h, w, _ = pages.shape
        pw = page_width_pt if page_width_pt is not None else w * 72.0 / self.cfg.dpi
        ph = page_height_pt if page_height_pt is not None else h * 72.0 / self.cfg.dpi
        n, stats, _ = self.detect_components(pages)
        n_h = n.cpu().numpy()
        nmax = int(np.abs(n_h).max()) if b else 0
        stats_h = stats[:, :max(nmax, 1)].cpu().numpy()
        if priors is not None and len(priors) != b:
            raise ValueError(f"priors: expected one entry per page ({b}), got {len(priors)}")
        per_page = [self.candidate_regions(stats_h[i], int(n_h[i]), pw, ph, priors[i] if priors is not None else None) for i in range(b)]
        rois, owners = [], []
        for i, regs in enumerate(per_page):
            for r in regs:
                x, y, cw, chh = self._crop_px(r["bbox"], w, h)
                r["crop_px"] = (x, y, cw, chh)
                r["page_num"] = int(page_nums[i]) if page_nums is not None else i
                rois.append((i, x, y, cw, chh))
                owners.append(r)
        if rois:
            mom = self.ctx.moments(pages, 1, rois).cpu().numpy()   # src_kind 1: moments of the PIL grey of each crop
            hashes = self.ctx.phash(pages, 1, rois).cpu().numpy() if with_hash else None
            for j, r in enumerate(owners):
                _, _, _, cw, chh = rois[j]
                npx = cw * chh
                s1, s2 = int(mom[j, 0]), int(mom[j, 1])
                var = (npx * s2 - s1 * s1) / (npx * npx)
                r["variance"] = var
                if r.get("detection_method") == "caption_based":
                    # regions handed in from the PDF object model are not re-scored: the reference keeps every caption
                    # region with confidence 0.9 (pdf_image_segmentation.py:2787-2799)
                    r.setdefault("confidence", 0.9)
                    r.setdefault("validation", "caption_based")
                else:
                    score, notes = G.validate_region(r["bbox"], cw, chh, var, ph)
                    r["confidence"] = score
                    r["validation"] = notes
                if hashes is not None:
                    r["phash"] = int(hashes[j]) & 0xFFFFFFFFFFFFFFFF
        out = []
        for regs in per_page:
            kept = [r for r in regs if r["confidence"] >= self.cfg.keep_score]
            kept.sort(key=lambda r: (r["bbox"].y0, r["bbox"].x0))
            out.append(kept)
        return out

    def detect_regions(self, page_rgb, page_num: int = 0, dpi: Optional[int] = None, page_width_pt: Optional[float] = None,
                       page_height_pt: Optional[float] = None, priors: Optional[Sequence[Dict]] = None) -> List[Dict]:
        """Single-page form mirroring `_detect_visual_regions(page, page_num) -> List[Dict]` (:3105)."""
        if dpi is not None and dpi != self.cfg.dpi:
            raise ValueError(f"detector configured for {self.cfg.dpi} DPI, got a {dpi} DPI page")
        t = page_rgb if isinstance(page_rgb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(page_rgb))
        t = t.to(self.ctx.device, non_blocking=True)
        return self.detect_regions_batch(t, [page_num], page_width_pt, page_height_pt, priors=[priors] if priors is not None else None)[0]

    def extract_segments(self, page_rgb: np.ndarray, page_num: int, book_id: str = "textbook_001",
                         output_dir: Optional[str] = None) -> List[VisualSegment]:
        """Mirror of `_extract_images_from_page(page, page_num) -> List[VisualSegment]` (:2763-2849) for a
        rasterised page: one VisualSegment per kept region with the reference's id scheme
        `{book}_p{page:03d}_{md5(png)[:8]}` (:3777-3783), page_no = page_num + 1, extraction_method =
        the region's detection method, confidence / notes from the validation score (:2915-2927)."""
        from PIL import Image
        regions = self.detect_regions(page_rgb, page_num)
        segs = []
        for r in regions:
            x, y, w, h = r["crop_px"]
            crop = Image.fromarray(np.ascontiguousarray(page_rgb[y:y + h, x:x + w]))
            segs.append(self.segment_from_region(r, crop, page_num, book_id, output_dir))
        return segs

    @staticmethod
    def segment_from_region(r: Dict, crop, page_num: int, book_id: str, output_dir: Optional[str] = None) -> VisualSegment:
        """Region dict + its PIL crop -> VisualSegment with the fields the reference sets at detection time
        (:2787-2799, :2915-2927); the crop is PNG-encoded for the id and, with output_dir, saved under it."""
        buf = io.BytesIO()
        crop.save(buf, format="PNG")
        png = buf.getvalue()
        seg_id = f"{book_id}_p{page_num:03d}_{hashlib.md5(png).hexdigest()[:8]}"
        path = None
        if output_dir is not None:
            import os
            os.makedirs(output_dir, exist_ok=True)
            path = os.path.join(output_dir, seg_id + ".png")
            with open(path, "wb") as f:
                f.write(png)
        return VisualSegment(segment_id=seg_id, segment_type=VisualType.UNKNOWN, book_id=book_id, page_no=page_num + 1,
                             bbox=r["bbox"], image_path=path, image_bytes=png, extraction_method=r["detection_method"],
                             caption_text=r["caption"], confidence=r["confidence"], notes=f"Validation: {r['validation']}")
