"""Raster counterpart of `VisualSegmentationPipeline.process()` (pdf_image_segmentation.py:2721-2761) for the part of
the flow that needs neither a PDF object model nor the network:

    rasterised pages -> GPU region detection (detector.py) -> one VisualSegment per kept region, crop saved as
    `{book}_p{page:03d}_{md5[:8]}.png` (:3777-3783) -> per-segment processing -> incremental JSON + summary CSV.

Per-segment processing is the reference's OFFLINE branch: with OCR and the Mistral API unavailable the reference
classifies every visual through `_fallback_analysis` (:701-715: type FIGURE, confidence 0.3, method
'fallback_heuristic') and, in the old algorithm, fills `FigureSpecificData` with the CV hints
(`process_figure_specific`, old_algo:986-1010) -- grid detection, arrow count > 3, grey variance > 1000 -- which here
run on the GPU (hints.py).  Like the reference (:2743-2754) a failing segment is reported and skipped, never fatal.
"""
from __future__ import annotations

import io
from typing import Iterable, List, Optional

import numpy as np
import torch
from PIL import Image

from .datamodel import OCRResult, VisualSegment, VisualType
from .detector import DetectConfig, RasterRegionDetector
from .hints import FeatureHints
from .writers import SegmentWriter

FALLBACK_SUMMARY = "Visual element detected (classification unavailable)"      # pdf_image_segmentation.py:712


class RasterSegmentationPipeline:
    def __init__(self, book_id: str, output_dir, dpi: int = 150, pdf_path: str = "", detector: Optional[RasterRegionDetector] = None,
                 batch: int = 8, with_hints: bool = True):
        self.book_id = book_id
        self.detector = detector or RasterRegionDetector(DetectConfig(dpi=dpi))
        self.writer = SegmentWriter(book_id, pdf_path, output_dir)
        self.batch = max(1, int(batch))
        self.with_hints = with_hints
        self.segments: List[VisualSegment] = []

    def _process_segment(self, seg: VisualSegment, crop: Image.Image) -> VisualSegment:
        """Offline `_process_segment` (:3659-3753 with the :701-715 fallback; hints as old_algo:3164-3183)."""
        seg.segment_type = VisualType.FIGURE
        seg.classification_confidence = 0.3
        seg.classification_method = "fallback_heuristic"
        seg.summary = FALLBACK_SUMMARY
        seg.summary_confidence = 0.3
        if self.with_hints:
            seg.ocr_result = OCRResult(detected_arrows=FeatureHints._count_arrows(crop))
            seg.figure_data = FeatureHints.process_figure_specific(crop, seg.ocr_result)
        return seg

    def process(self, pages: Iterable[np.ndarray], first_page: int = 0) -> List[VisualSegment]:
        """pages: iterable of HxWx3 u8 page rasters at the detector's DPI.  Returns all segments (also written)."""
        self.writer.initialize()
        buf: List[np.ndarray] = []
        page_num = first_page
        try:
            for page in pages:
                page = np.ascontiguousarray(page)
                if buf and page.shape != buf[0].shape:      # a batch holds pages of ONE size (PDFs mix page sizes)
                    self._process_batch(buf, page_num)
                    page_num += len(buf)
                    buf = []
                buf.append(page)
                if len(buf) == self.batch:
                    self._process_batch(buf, page_num)
                    page_num += len(buf)
                    buf = []
            if buf:
                self._process_batch(buf, page_num)
        finally:
            self.writer.save_results()            # the reference saves in `finally` as well (:2755-2758)
        return self.segments

    def _process_batch(self, pages: List[np.ndarray], first_page: int) -> None:
        det = self.detector
        try:
            batch = torch.from_numpy(np.stack(pages)).to(det.ctx.device, non_blocking=True)
            # pass 2 of _extract_images_from_page runs on every page (no caption-based priors without a PDF object model)
            regions = det.detect_regions_batch(batch, page_nums=list(range(first_page, first_page + len(pages))), priors=[[]] * len(pages))
        except Exception as e:                      # a failing batch is reported and skipped like a failing segment (:2749-2754)
            print(f"    ERROR detecting regions on pages {first_page + 1}..{first_page + len(pages)}: {e}")
            return
        for i, regs in enumerate(regions):
            page_num = first_page + i
            for r in regs:
                try:
                    x, y, w, h = r["crop_px"]
                    crop = Image.fromarray(np.ascontiguousarray(pages[i][y:y + h, x:x + w]))
                    seg = det.segment_from_region(r, crop, page_num, self.book_id, str(self.writer.output_dir))
                    seg = self._process_segment(seg, crop)
                    if self.writer.append(seg):
                        self.segments.append(seg)
                except Exception as e:              # per-segment try/except like :2749-2754
                    print(f"    ERROR processing a segment of page {page_num + 1}: {e}")
                    continue
