"""Pinned, multi-buffered host->device page streaming: the rasterisation handoff of this stage.

The reference hands a rasterised region over as RGB u8 pixels (pdf_image_segmentation.py:3638-3657).  Here
whole pages arrive in pinned host memory (row-major HWC, stride 3W, no alpha); a copy stream moves batches
i+1, i+2 to the device while the compute stream runs the fused detection pipeline on batch i, and the small
result tensors (n_labels, stats) come back through pinned buffers.  The host consumes the results of batch
i-slots+1 (box filter / merge in Python) only after the next copy has been queued, so the copy engine never
waits for Python.  End to end this stage is PCIe-bound (25.2 MB per 300-DPI page), not HBM-bound.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import numpy as np
import torch

from .detector import RasterRegionDetector


class PageStreamer:
    def __init__(self, detector: RasterRegionDetector, batch: int, height: int, width: int, slots: int = 3):
        self.det = detector
        self.batch, self.h, self.w = batch, height, width
        dev = detector.ctx.device
        ml = detector.cfg.max_labels
        self.slots = slots
        self.dev_pages = [torch.empty((batch, height, width, 3), dtype=torch.uint8, device=dev) for _ in range(slots)]
        self.dev_out = [(torch.empty(batch, dtype=torch.int32, device=dev),
                         torch.empty((batch, ml, 5), dtype=torch.int32, device=dev),
                         torch.empty((batch, ml, 2), dtype=torch.float64, device=dev)) for _ in range(slots)]
        self.host_n = [torch.empty(batch, dtype=torch.int32).pin_memory() for _ in range(slots)]
        self.host_stats = [torch.empty((batch, ml, 5), dtype=torch.int32).pin_memory() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in range(slots)]
        self.computed = [torch.cuda.Event() for _ in range(slots)]
        self.drained = [torch.cuda.Event() for _ in range(slots)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        detector.ctx.reserve(width, height, batch)

    def run(self, host_batches: Iterable[torch.Tensor], on_result: Optional[Callable] = None) -> int:
        """host_batches: pinned u8 tensors [b<=batch, H, W, 3].  Calls on_result(batch_index, n_labels, stats)
        with host tensors (valid until the slot is reused).  Returns the number of pages processed."""
        pending: List[tuple] = []
        pages = 0
        for i, hb in enumerate(host_batches):
            s = i % self.slots
            b = hb.shape[0]
            # the slot's previous batch must be consumed (host-side wait on its `computed` event) before its
            # page buffer is overwritten and its result buffers are reused; up to slots-1 batches stay queued
            # on the device while the host post-processes the oldest one
            while len(pending) >= self.slots:
                self._finish(pending.pop(0), on_result)
            with torch.cuda.stream(self.copy_stream):
                self.dev_pages[s][:b].copy_(hb, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            self.h2d_bytes += hb.numel()
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[s])
                n, st, ce = self.dev_out[s]
                self.det.detect_components(self.dev_pages[s][:b], out=(n[:b], st[:b], ce[:b]))
                # contiguous -> contiguous pinned copies only: a strided device->host copy_ goes through a staging
                # buffer and BLOCKS the host, which would serialise the next H2D behind this batch's compute
                self.host_n[s][:b].copy_(n[:b], non_blocking=True)
                self.host_stats[s][:b].copy_(st[:b], non_blocking=True)
                self.computed[s].record(self.compute_stream)
            self.d2h_bytes += b * 4 + b * self.host_stats[s].shape[1] * 20
            pending.append((i, s, b))
            pages += b
        while pending:
            self._finish(pending.pop(0), on_result)
        return pages

    def _finish(self, item, on_result):
        i, s, b = item
        self.computed[s].synchronize()
        if on_result is not None:
            on_result(i, self.host_n[s][:b], self.host_stats[s][:b])
