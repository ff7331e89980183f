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
    max_labels: int = 1024               # component-table rows per page; a page with more is re-run with a larger table
    max_regions: int = 64                # region-list capacity per page on the device (the shipped run has <= 14 per page)
    min_extent_pt: float = 50.0          # :3450 (drawing / image regions must exceed 50 x 50 pt)
    keep_score: float = G.KEEP_SCORE     # :2885

    def resolved(self):
        bs = self.block_size or ((self.dpi // 6) | 1)
        k = self.k or (int(10 * self.dpi / 72) | 1)
        return bs, self.C, k


_contexts: Dict[int, Context] = {}


def get_context(device: Optional[int] = None) -> Context:
    """The process-wide context of a CUDA device (default: torch's current device); one context per (process, GPU)."""
    if device is None:
        device = torch.cuda.current_device()
    ctx = _contexts.get(device)
    if ctx is None or ctx._h is None:
        ctx = _contexts[device] = Context(device)
    return ctx


class RasterRegionDetector:
    def __init__(self, config: Optional[DetectConfig] = None, ctx: Optional[Context] = None, device: Optional[int] = None):
        self.cfg = config or DetectConfig()
        self.ctx = ctx or get_context(device)

    # ---- GPU stage ---------------------------------------------------------------------------
    def detect_components(self, pages: torch.Tensor, out=None, max_labels: Optional[int] = None):
        """pages: CUDA u8 [B,H,W,3] RGB (or [H,W,3]) or [B,H,W] grey.  Returns device tensors (n_labels, stats, centroids)."""
        bs, c, k = self.cfg.resolved()
        return self.ctx.detect_pages(pages, bs, c, k, self.cfg.canny_lo, self.cfg.canny_hi, max_labels or self.cfg.max_labels, out=out)

    def detect_tables(self, pages: torch.Tensor, page_width_pt: float, page_height_pt: float, out=None):
        """Device pages -> dict of device tensors (n_labels, stats, regions, n_regions, flags): component tables AND the
        candidate regions with their crop moments, computed in one stream (synseg_detect_regions)."""
        bs, c, k = self.cfg.resolved()
        return self.ctx.detect_regions(pages, bs, c, k, self.cfg.dpi, page_width_pt, page_height_pt, self.cfg.canny_lo, self.cfg.canny_hi,
                                       self.cfg.max_labels, self.cfg.max_regions, self.cfg.min_extent_pt, out=out)

    # ---- host stage --------------------------------------------------------------------------
    def raster_regions(self, stats: np.ndarray, n_labels: int, page_width_pt: float, page_height_pt: float) -> List[Dict]:
        """Component stats of one page ([n,5] = x,y,w,h,area in px; row 0 = background) -> raster region dicts (bbox in
        points) before validation: the host form of csrc/regions.cu (which is bit-identical and is what normally runs)."""
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
        return G.merge_visual_regions(primary, secondary)

    def candidate_regions(self, stats: np.ndarray, n_labels: int, page_width_pt: float, page_height_pt: float,
                          priors: Optional[Sequence[Dict]] = None) -> List[Dict]:
        """Raster regions of one page merged with PDF-structural priors by `_detect_visual_regions`' rule.

        priors: regions a PDF object model produced for this page (the reference's caption-based regions,
        `_detect_by_captions`, pdf_image_segmentation.py:3148-3254: dicts with 'bbox' and, when known, 'caption_bbox').
        They take the place `caption_regions` has in `_detect_visual_regions` (:3114-3144): every prior is kept, a raster
        region is added unless more than half of it lies inside a prior or it sits right above a prior's caption
        (SURVEY.md 8f rank 4)."""
        return self.apply_priors_visual_regions(self.raster_regions(stats, n_labels, page_width_pt, page_height_pt), priors)

    @staticmethod
    def apply_priors_visual_regions(raster: List[Dict], priors: Optional[Sequence[Dict]]) -> List[Dict]:
        if not priors:
            return raster
        # every raster region is tested against the PRIORS only (among themselves the raster regions were merged already:
        # with no priors the result must not change), by the two rules of _detect_visual_regions
        priors = list(priors)
        return priors + [r for r in raster
                         if not G.overlaps_with_existing(r["bbox"], priors)
                         and not any("caption_bbox" in p and G.caption_near_region(r["bbox"], p["caption_bbox"]) for p in priors)]

    def regions_from_table(self, table: np.ndarray, n_regions: int, page_width_pt: float, page_height_pt: float) -> List[Dict]:
        """One page of the device's region table (REGION_DTYPE rows) -> region dicts identical to `raster_regions` + crop + moments."""
        out = []
        for r in table[:n_regions].tolist():
            x0, y0, x1, y1, px, py, pw, ph, kind, count, s1, s2 = r
            if kind == 2:
                # the host rule clamps with max(0, .): Python hands back the int 0 there (it shows in the JSON as 0, not 0.0)
                x0 = 0 if x0 == 0.0 else x0
                y0 = 0 if y0 == 0.0 else y0
                notes, method = f"Detected from {count} raster components", "raster_cluster"
            else:
                notes, method = f"Connected component of {count} px", "raster_cc"
            out.append({"bbox": BoundingBox(x0, y0, x1, y1, page_width_pt, page_height_pt), "caption": None, "detection_method": method,
                        "notes": notes, "crop_px": (px, py, pw, ph), "_moments": (int(s1), int(s2))})
        return out

    def _crop_px(self, bbox: BoundingBox, width: int, height: int):
        x, y, w, h = bbox.to_pixels(self.cfg.dpi)
        x = min(max(x, 0), width - 1); y = min(max(y, 0), height - 1)
        w = max(1, min(w, width - x)); h = max(1, min(h, height - y))
        return x, y, w, h

    def _score(self, regs: List[Dict], page_height_pt: float) -> None:
        """variance from the exact moments, then the _validate_embedded_image score (:2933-2998) -- in place."""
        for r in regs:
            _, _, cw, chh = r["crop_px"]
            npx = cw * chh
            s1, s2 = r.pop("_moments")
            var = (npx * s2 - s1 * s1) / (npx * npx)
            r["variance"] = var
            if r.get("detection_method") == "caption_based":
                # regions handed in from the PDF object model are not re-scored: the reference keeps every caption
                # region with confidence 0.9 (pdf_image_segmentation.py:2787-2799)
                r.setdefault("confidence", 0.9)
                r.setdefault("validation", "caption_based")
            else:
                score, notes = G.validate_region(r["bbox"], cw, chh, var, page_height_pt)
                r["confidence"] = score
                r["validation"] = notes

    def _host_moments(self, page: torch.Tensor, regs: List[Dict], width: int, height: int) -> None:
        """crop_px + moments for region dicts that did not come from the device table (host fallback, priors)."""
        todo = [r for r in regs if "_moments" not in r]
        if not todo:
            return
        for r in todo:
            r["crop_px"] = self._crop_px(r["bbox"], width, height)
        grey = page.dim() == 2
        mom = self.ctx.moments(page if not grey else page[None], 0 if grey else 1, [(0,) + r["crop_px"] for r in todo]).cpu().numpy()
        for r, m in zip(todo, mom):
            r["_moments"] = (int(m[0]), int(m[1]))

    def _page_fallback(self, page: torch.Tensor, n_labels: int, stats_row: Optional[np.ndarray], pw: float, ph: float) -> List[Dict]:
        """A page the device could not decide (flags != 0): more components than max_labels -> re-run the page with a table
        that holds them all; then the host rules on the component table."""
        if n_labels < 0 or stats_row is None:
            need = max(-n_labels + 1, self.cfg.max_labels)
            n, stats, _ = self.detect_components(page[None], max_labels=need)
            n_labels = int(n[0])
            if n_labels < 0:                          # cannot happen: the table now holds the count reported before
                raise RuntimeError(f"page has {-n_labels} components, more than the retry capacity {need}")
            stats_row = stats[0, :n_labels].cpu().numpy()
        return self.raster_regions(stats_row, n_labels, pw, ph)

    def finish_regions(self, tables: Dict[str, np.ndarray], pages, page_nums: Optional[Sequence[int]], pw: float, ph: float,
                       width: int, height: int, with_hash: bool = False, priors=None, prior_rule: str = "two_pass",
                       drawings=None) -> List[List[Dict]]:
        """Host end of the stage: region tables of a batch (numpy views: n_labels, stats | None, regions, n_regions, flags) ->
        per page the validated, filtered, sorted region dicts.  `pages` (device tensor, or a callable page index -> device page)
        is only touched for flagged pages, priors and hashes."""
        b = tables["n_regions"].shape[0]
        get_page = pages if callable(pages) else (lambda i: pages[i])
        out: List[List[Dict]] = []
        for i in range(b):
            if int(tables["flags"][i]):
                nl = int(tables["n_labels"][i])
                st_row = tables["stats"][i, :nl] if (tables.get("stats") is not None and nl > 0) else None
                raster = self._page_fallback(get_page(i), nl, st_row, pw, ph)
            else:
                raster = self.regions_from_table(tables["regions"][i], int(tables["n_regions"][i]), pw, ph)
            pr = priors[i] if priors is not None else None
            if pr is not None:
                pr = [dict(p, detection_method=p.get("detection_method", "caption_based"), caption=p.get("caption")) for p in pr]
            if pr and prior_rule == "visual_regions":
                regs = self.apply_priors_visual_regions(raster, pr)
            else:
                regs = raster + [p for p in (pr or [])]
            if any("_moments" not in r for r in regs):
                self._host_moments(get_page(i), regs, width, height)
            self._score(regs, ph)
            for r in regs:
                r["page_num"] = int(page_nums[i]) if page_nums is not None else i
            kept = [r for r in regs if r["confidence"] >= self.cfg.keep_score]
            if pr is not None and prior_rule == "two_pass":
                # _extract_images_from_page (:2763-2849): caption-based regions are pass 1 (possibly none); every validated raster
                # region is a pass-2 candidate resolved against the segments kept so far (find_conflicting > 0.4, five-factor
                # vote) -- also against candidates added before it, exactly like overlapping embedded images in the reference
                caps = [r for r in kept if r.get("detection_method") == "caption_based"]
                cands = [r for r in kept if r.get("detection_method") != "caption_based"]
                kept = G.resolve_page_conflicts(caps, cands, drawings[i] if drawings is not None else None)
            if with_hash and kept:
                page = get_page(i)
                grey = page.dim() == 2
                hs = self.ctx.phash(page[None] if grey else page, 0 if grey else 1, [(0,) + r["crop_px"] for r in kept]).cpu().numpy()
                for r, hv in zip(kept, hs):
                    r["phash"] = int(hv) & 0xFFFFFFFFFFFFFFFF
            kept.sort(key=lambda r: (r["bbox"].y0, r["bbox"].x0))
            out.append(kept)
        return out

    def detect_regions_batch(self, pages: torch.Tensor, page_nums: Optional[Sequence[int]] = None,
                             page_width_pt: Optional[float] = None, page_height_pt: Optional[float] = None,
                             with_hash: bool = False, priors: Optional[Sequence[Optional[Sequence[Dict]]]] = None,
                             prior_rule: str = "two_pass", drawings: Optional[Sequence[Optional[Sequence[Sequence[float]]]]] = None) -> List[List[Dict]]:
        """Pages (CUDA u8 [B,H,W,3] RGB or [B,H,W] grey) -> per page a list of region dicts, sorted by (y0, x0).

        Region dict = the reference's schema (pdf_image_segmentation.py:3246-3252 / 3550-3555):
        {'bbox': BoundingBox (points), 'caption': None, 'detection_method', 'notes'} plus
        'confidence' / 'validation' (the _validate_embedded_image score and notes), 'variance', 'page_num',
        'crop_px' (x, y, w, h) and, with_hash, 'phash'.

        priors: per page the regions a PDF object model produced (caption-based regions with 'bbox', 'caption', optional
        'caption_bbox'; detection_method 'caption_based'); an EMPTY list for a page still runs pass 2 among the raster
        candidates of that page, None (or priors=None) returns the validated regions as detected.
        prior_rule "two_pass" (default) follows the live flow of the
        reference, `_extract_images_from_page` pass 2 (:2822-2847): validated raster regions are candidates resolved against
        the caption-based segments by `_find_conflicting_segment` / `_resolve_conflict` (`drawings`: per page the drawing
        rects for factor 4).  "visual_regions" applies `_detect_visual_regions`' duplicate / caption rule (:3122-3144) instead."""
        if prior_rule not in ("two_pass", "visual_regions"):
            raise ValueError("prior_rule must be 'two_pass' or 'visual_regions'")
        rgb = pages.dim() == 4 or (pages.dim() == 3 and pages.shape[-1] == 3 and pages.stride(-2) == 3)
        if pages.dim() == (3 if rgb else 2):
            pages = pages.unsqueeze(0)
        b, h, w = pages.shape[0], pages.shape[1], pages.shape[2]
        pw = page_width_pt if page_width_pt is not None else w * 72.0 / self.cfg.dpi
        ph = page_height_pt if page_height_pt is not None else h * 72.0 / self.cfg.dpi
        if priors is not None and len(priors) != b:
            raise ValueError(f"priors: expected one entry per page ({b}), got {len(priors)}")
        if priors is not None:
            priors = [list(pp) if pp is not None else None for pp in priors]
        t = self.detect_tables(pages, pw, ph)
        n_h = t["n_labels"].cpu().numpy()
        tables = dict(n_labels=n_h, n_regions=t["n_regions"].cpu().numpy(), flags=t["flags"].cpu().numpy(),
                      regions=Context.regions_view(t["regions"].cpu()), stats=None)
        if tables["flags"].any():
            nmax = int(np.abs(n_h).max())
            tables["stats"] = t["stats"][:, :max(min(nmax, self.cfg.max_labels), 1)].cpu().numpy()
        return self.finish_regions(tables, pages, page_nums, pw, ph, w, h, with_hash, priors, prior_rule, drawings)

    def detect_regions(self, page_rgb, page_num: int = 0, dpi: Optional[int] = None, page_width_pt: Optional[float] = None,
                       page_height_pt: Optional[float] = None, priors: Optional[Sequence[Dict]] = None, prior_rule: str = "two_pass",
                       drawings: Optional[Sequence[Sequence[float]]] = None) -> List[Dict]:
        """Single-page form mirroring `_detect_visual_regions(page, page_num) -> List[Dict]` (:3105)."""
        if dpi is not None and dpi != self.cfg.dpi:
            raise ValueError(f"detector configured for {self.cfg.dpi} DPI, got a {dpi} DPI page")
        t = page_rgb if isinstance(page_rgb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(page_rgb))
        t = t.to(self.ctx.device, non_blocking=True)
        return self.detect_regions_batch(t, [page_num], page_width_pt, page_height_pt, priors=[priors] if priors is not None else None,
                                         prior_rule=prior_rule, drawings=[drawings] if drawings is not None else None)[0]

    def extract_segments(self, page_rgb: np.ndarray, page_num: int, book_id: str = "textbook_001",
                         output_dir: Optional[str] = None, priors: Optional[Sequence[Dict]] = None,
                         drawings: Optional[Sequence[Sequence[float]]] = None) -> List[VisualSegment]:
        """Mirror of `_extract_images_from_page(page, page_num) -> List[VisualSegment]` (:2763-2849) for a
        rasterised page: one VisualSegment per kept region with the reference's id scheme
        `{book}_p{page:03d}_{md5(png)[:8]}` (:3777-3783), page_no = page_num + 1, extraction_method =
        the region's detection method, confidence / notes from the validation score (:2915-2927)."""
        from PIL import Image
        regions = self.detect_regions(page_rgb, page_num, priors=priors if priors is not None else [], prior_rule="two_pass", drawings=drawings)
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
        caption_based = r["detection_method"] == "caption_based"        # pass-1 segments keep their own notes (:2787-2799)
        return VisualSegment(segment_id=seg_id, segment_type=VisualType.UNKNOWN, book_id=book_id, page_no=page_num + 1,
                             bbox=r["bbox"], image_path=path, image_bytes=png, extraction_method=r["detection_method"],
                             caption_text=r.get("caption"), confidence=r["confidence"],
                             notes=r.get("notes", "") if caption_based else f"Validation: {r['validation']}")
