"""Pinned, multi-buffered host->device page streaming: the rasterisation handoff of this stage.

The reference hands a rasterised region over as RGB (or L) u8 pixels (pdf_image_segmentation.py:3638-3657).  Here
whole pages arrive in pinned host memory (row-major; RGB interleaved [N,H,W,3] or grey [N,H,W]); the library
(`synseg_detect_regions_host`) stages them chunk by chunk through a three-slot device ring on its own copy stream,
runs detection, the region rules and the crop moments on the compute stream and returns, per page, the candidate
regions with their exact grey moments plus the component table -- so what reaches Python is a few hundred bytes per
page that only need the `_validate_embedded_image` score (f64 adds in the reference's order) and the keep >= 0.5
filter.  The host finishes batch i-slots+1 only after batch i has been queued, so the copy engine never waits for
Python.  End to end this stage is PCIe-bound (25.2 MB per 300-DPI RGB page, 8.4 MB per grey page), not HBM-bound.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch

from .detector import RasterRegionDetector
from .ops import Context


class PageStreamer:
    def __init__(self, detector: RasterRegionDetector, batch: int, height: int, width: int, slots: int = 3, chunk_pages: int = 10,
                 channels: int = 3, page_width_pt: Optional[float] = None, page_height_pt: Optional[float] = None):
        self.det = detector
        self.batch, self.h, self.w, self.channels = batch, height, width, channels
        cfg = detector.cfg
        self.pw = page_width_pt if page_width_pt is not None else width * 72.0 / cfg.dpi
        self.ph = page_height_pt if page_height_pt is not None else height * 72.0 / cfg.dpi
        self.slots = slots
        self.chunk_pages = max(1, min(chunk_pages, batch))
        self.out = [Context.region_buffers(batch, cfg.max_labels, cfg.max_regions, pinned=True) for _ in range(slots)]
        self.stream = torch.cuda.Stream(device=detector.ctx.device)
        self.finished = [torch.cuda.Event() for _ in range(slots)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        detector.ctx.reserve(width, height, self.chunk_pages)

    def run(self, host_batches: Iterable[torch.Tensor], on_result: Optional[Callable] = None, on_regions: Optional[Callable] = None,
            page_base: int = 0) -> int:
        """host_batches: pinned u8 tensors [b<=batch, H, W, 3] (RGB) or [b, H, W] (grey).
        on_result(batch_index, n_labels, stats): the raw component tables (host tensors, valid until the slot is reused);
        on_regions(batch_index, regions): per page the validated, filtered region dicts of `RasterRegionDetector.detect_regions_batch`
        (same dicts, same order).  Returns the number of pages processed."""
        det, cfg = self.det, self.det.cfg
        bs, c, k = cfg.resolved()
        pending: List[tuple] = []
        pages = 0
        for i, hb in enumerate(host_batches):
            s = i % self.slots
            b = hb.shape[0]
            # the slot's previous batch must be consumed before its result buffers are reused; up to slots-1 batches stay
            # queued on the device while the host post-processes the oldest one
            while len(pending) >= self.slots:
                self._finish(pending.pop(0), on_result, on_regions, page_base)
            o = {name: t[:b] for name, t in self.out[s].items()}
            with torch.cuda.stream(self.stream):
                det.ctx.detect_regions_host(hb, bs, c, k, cfg.dpi, self.pw, self.ph, cfg.canny_lo, cfg.canny_hi, cfg.max_labels, cfg.max_regions,
                                            cfg.min_extent_pt, chunk_pages=self.chunk_pages, out=o)
                self.finished[s].record(self.stream)
            self.h2d_bytes += hb.numel()
            self.d2h_bytes += sum(t.numel() * t.element_size() for t in o.values())
            pending.append((i, s, b, hb, pages))
            pages += b
        while pending:
            self._finish(pending.pop(0), on_result, on_regions, page_base)
        return pages

    def _finish(self, item, on_result, on_regions, page_base):
        i, s, b, hb, first = item
        self.finished[s].synchronize()
        o = self.out[s]
        if on_result is not None:
            on_result(i, o["n_labels"][:b], o["stats"][:b])
        if on_regions is not None:
            det = self.det
            tables = dict(n_labels=o["n_labels"][:b].numpy(), stats=o["stats"][:b].numpy(), n_regions=o["n_regions"][:b].numpy(),
                          flags=o["flags"][:b].numpy(), regions=Context.regions_view(o["regions"][:b]))
            # flagged pages (rare) are re-uploaded from the host batch: their ring slot on the device has been reused
            get_page = lambda j: hb[j].to(det.ctx.device)        # noqa: E731
            on_regions(i, det.finish_regions(tables, get_page, range(page_base + first, page_base + first + b), self.pw, self.ph, self.w, self.h))
