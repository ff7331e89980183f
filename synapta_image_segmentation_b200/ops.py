"""Tensor-facing wrappers of the libsynseg C ABI.

PyTorch is used for device memory, streams and nothing else: every function takes/returns CUDA
``torch.uint8`` (or int32/uint64-as-int64) tensors, builds ``synseg_img`` descriptors from their
pointers and strides and calls the C ABI on the current CUDA stream.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Crop, DetectParams, Img, Region, RegionParams, Roi, check

REGION_DTYPE = np.dtype([("x0", "<f8"), ("y0", "<f8"), ("x1", "<f8"), ("y1", "<f8"), ("px", "<i4"), ("py", "<i4"), ("pw", "<i4"), ("ph", "<i4"),
                         ("kind", "<i4"), ("count", "<i4"), ("sum", "<u8"), ("sum_sq", "<u8")])     # synseg_region
assert REGION_DTYPE.itemsize == _lib.REGION_BYTES == C.sizeof(Region)

ERODE, DILATE, OPEN, CLOSE = 0, 1, 2, 3
GRAY_CV, GRAY_PIL = 0, 1
HIST_BINS = 4096


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as3(t: torch.Tensor, channels: int) -> torch.Tensor:
    """Normalise to [B,H,W] (grey) or [B,H,W,3] (RGB) without copying."""
    want = 3 if channels == 1 else 4
    if t.dim() == want - 1:
        t = t.unsqueeze(0)
    if t.dim() != want:
        raise ValueError(f"expected a {want - 1}- or {want}-d tensor, got shape {tuple(t.shape)}")
    return t


def img_of(t: torch.Tensor, channels: int = 1, itemsize: int = 1) -> Img:
    """synseg_img descriptor of a CUDA tensor ([H,W] / [B,H,W] grey, [H,W,3] / [B,H,W,3] RGB)."""
    if not t.is_cuda:
        raise ValueError("libsynseg operates on CUDA tensors only (no CPU fallback)")
    if t.element_size() != itemsize:
        raise ValueError(f"expected {itemsize}-byte elements, got {t.dtype}")
    t = _as3(t, channels)
    if channels == 3:
        if t.shape[-1] != 3 or t.stride(-1) != 1 or t.stride(-2) != 3:
            raise ValueError("RGB images must be interleaved HWC with contiguous pixels")
        b, h, w, _ = t.shape
        rs, bs = t.stride(1), t.stride(0)
    else:
        if t.stride(-1) != 1:
            raise ValueError("rows must be contiguous")
        b, h, w = t.shape
        rs, bs = t.stride(1) * itemsize, t.stride(0) * itemsize
    if h == 0 or w == 0 or b == 0:
        raise ValueError("empty image")
    if b == 1:
        bs = rs * h
    return Img(t.data_ptr(), w, h, rs, b, 0, bs)


def empty_plane(batch: int, height: int, width: int, device, dtype=torch.uint8, pitch_align: int = 128) -> torch.Tensor:
    """[B,H,W] view into a pitched allocation whose rows are `pitch_align`-byte aligned."""
    isz = torch.empty((), dtype=dtype).element_size()
    pitch = (width * isz + pitch_align - 1) // pitch_align * pitch_align // isz
    return torch.empty((batch, height, pitch), dtype=dtype, device=device)[:, :, :width]


def rois_array(rois: Sequence[Sequence[int]]) -> np.ndarray:
    """[(image, x, y, w, h), ...] -> int32 [n,5] array laid out as synseg_roi[]."""
    a = np.ascontiguousarray(np.asarray(rois, dtype=np.int32).reshape(-1, 5))
    return a


class Context:
    """One synseg context per (process, GPU).  Not thread-safe."""

    def __init__(self, device: int | torch.device | None = None):
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        h = C.c_void_p()
        check(self.lib.synseg_create(self.device.index, C.byref(h)), "synseg_create")
        self._h = h

    def _s(self) -> C.c_void_p:
        """The current torch stream of THIS context's device (not of torch's current device)."""
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _img(self, t: torch.Tensor, channels: int = 1, itemsize: int = 1) -> Img:
        if t.is_cuda and t.device != self.device:
            raise ValueError(f"tensor lives on {t.device}, this context on {self.device}")
        return img_of(t, channels, itemsize)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.synseg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.synseg_launch_count(self._h))

    def reserve(self, width: int, height: int, batch: int):
        check(self.lib.synseg_reserve(self._h, self.lib.synseg_scratch_bytes(width, height, batch)), "synseg_reserve")

    # ---- CUDA graphs -----------------------------------------------------------------------------
    def capture(self, fn, warmup: int = 2):
        """Captures the library calls `fn()` makes on this context's device into a CUDA graph and returns `replay()`.
        One graph launch instead of ~45 kernel launches per 50-page step: the launch-bound inner loop of a resident pipeline
        (at 8 ranks on one host the per-launch cost of the driver, not the GPUs, limits the step otherwise).  `fn` must use
        fixed tensors (the graph bakes pointers and shapes in) and must not synchronise; it runs `warmup` times first so
        that scratch, side streams and kernel attributes exist before the capture.  The caller orders replays against other
        work on the context (calls inside a capture skip the context's cross-stream event, see csrc/internal.cuh)."""
        for _ in range(max(1, warmup)):
            fn()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        with torch.cuda.device(self.device), torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            fn()
        return g.replay

    # ---- per-kernel timing ---------------------------------------------------------------------
    def profile_begin(self):
        check(self.lib.synseg_profile_begin(self._h, self._s()), "synseg_profile_begin")

    def profile_end(self, cap: int = 4096):
        """[(kernel name, milliseconds), ...] in launch order since profile_begin()."""
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = self.lib.synseg_profile_end(self._h, names, ms, cap)
        if n < 0:
            check(n, "synseg_profile_end")
        return [(names[i].decode(), float(ms[i])) for i in range(min(n, cap))]

    # ---- colour -------------------------------------------------------------------------------
    def rgb2gray(self, rgb: torch.Tensor, mode: int = GRAY_CV, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        r = _as3(rgb, 3)
        if out is None:
            out = empty_plane(r.shape[0], r.shape[1], r.shape[2], rgb.device)
        check(self.lib.synseg_rgb2gray(self._h, C.byref(self._img(rgb, 3)), C.byref(self._img(out)), mode, self._s()), "synseg_rgb2gray")
        return out if rgb.dim() == 4 else out[0]

    # ---- threshold / edges ----------------------------------------------------------------------
    def adaptive_mean(self, gray: torch.Tensor, block_size: int, c: int, invert: bool = True,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = _as3(gray, 1)
        if out is None:
            out = empty_plane(*g.shape, gray.device)
        check(self.lib.synseg_adaptive_mean(self._h, C.byref(self._img(gray)), C.byref(self._img(out)), block_size, c, int(invert), self._s()),
              "synseg_adaptive_mean")
        return out if gray.dim() == 3 else out[0]

    def canny(self, gray: torch.Tensor, lo: int = 50, hi: int = 150, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = _as3(gray, 1)
        if out is None:
            out = empty_plane(*g.shape, gray.device)
        check(self.lib.synseg_canny(self._h, C.byref(self._img(gray)), C.byref(self._img(out)), lo, hi, self._s()), "synseg_canny")
        return out if gray.dim() == 3 else out[0]

    # ---- morphology ---------------------------------------------------------------------------
    def morph(self, src: torch.Tensor, op: int, kw: int, kh: int, anchor=(-1, -1), iterations: int = 1, binary: bool = False,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        s = _as3(src, 1)
        if out is None:
            out = empty_plane(*s.shape, src.device)
        check(self.lib.synseg_morph(self._h, C.byref(self._img(src)), C.byref(self._img(out)), op, kw, kh, anchor[0], anchor[1], iterations,
                                    1 if binary else 0, self._s()), "synseg_morph")
        return out if src.dim() == 3 else out[0]

    # ---- connected components -----------------------------------------------------------------
    def ccl_stats(self, mask: torch.Tensor, max_labels: int = 4096, want_labels: bool = True):
        """Returns (n_labels int32 [B], labels int32 [B,H,W] | None, stats int32 [B,max,5], centroids f64 [B,max,2])."""
        m = _as3(mask, 1)
        b, h, w = m.shape
        dev = mask.device
        labels = empty_plane(b, h, w, dev, torch.int32) if want_labels else None
        n = torch.empty(b, dtype=torch.int32, device=dev)
        stats = torch.empty((b, max_labels, 5), dtype=torch.int32, device=dev)
        cent = torch.empty((b, max_labels, 2), dtype=torch.float64, device=dev)
        limg = C.byref(self._img(labels, 1, 4)) if want_labels else None
        check(self.lib.synseg_ccl_stats(self._h, C.byref(self._img(mask)), limg, n.data_ptr(), stats.data_ptr(), cent.data_ptr(),
                                        max_labels, self._s()), "synseg_ccl_stats")
        return n, labels, stats, cent

    # ---- reductions ---------------------------------------------------------------------------
    def _rois_dev(self, rois, device):
        if rois is None:
            return None, 0, None
        a = rois_array(rois)
        t = torch.from_numpy(a).to(device, non_blocking=False)
        return t, a.shape[0], a

    @staticmethod
    def _check_rois(a: np.ndarray, b: int, h: int, w: int):
        if a is None:
            return
        if ((a[:, 0] < 0) | (a[:, 0] >= b) | (a[:, 1] < 0) | (a[:, 2] < 0) | (a[:, 3] <= 0) | (a[:, 4] <= 0)
                | (a[:, 1] + a[:, 3] > w) | (a[:, 2] + a[:, 4] > h)).any():
            raise ValueError("region outside the image")

    def moments(self, src: torch.Tensor, src_kind: int = 0, rois=None) -> torch.Tensor:
        """int64 [n,3] = (sum, sum of squares, non-zero count) per region (whole images if rois is None)."""
        ch = 3 if src_kind else 1
        s = _as3(src, ch)
        rt, n, a = self._rois_dev(rois, src.device)
        self._check_rois(a, s.shape[0], s.shape[1], s.shape[2])
        if rois is None:
            n = s.shape[0]
        out = torch.empty((n, 3), dtype=torch.int64, device=src.device)
        if n:
            check(self.lib.synseg_moments(self._h, C.byref(self._img(src, ch)), src_kind, rt.data_ptr() if rt is not None else None, n,
                                          out.data_ptr(), self._s()), "synseg_moments")
        return out

    def hsv_mask_hist(self, rgb: torch.Tensor, rois=None, want_hist: bool = True, want_sums: bool = False, want_rows: bool = False):
        """Returns dict(count int64 [n], hist int32 [n,4096] | None, chan_sum int64 [n,4096,3] | None, row_count int32 [n,max_rows] | None)."""
        s = _as3(rgb, 3)
        rt, n, a = self._rois_dev(rois, rgb.device)
        self._check_rois(a, s.shape[0], s.shape[1], s.shape[2])
        if rois is None:
            n = s.shape[0]
            max_rows = s.shape[1]
        else:
            max_rows = int(a[:, 4].max()) if n else 0
        dev = rgb.device
        count = torch.empty(n, dtype=torch.int64, device=dev)
        hist = torch.empty((n, HIST_BINS), dtype=torch.int32, device=dev) if (want_hist or want_sums) else None
        sums = torch.empty((n, HIST_BINS, 3), dtype=torch.int64, device=dev) if want_sums else None
        rows = torch.empty((n, max_rows), dtype=torch.int32, device=dev) if want_rows else None
        if n:
            check(self.lib.synseg_hsv_mask_hist(self._h, C.byref(self._img(rgb, 3)), rt.data_ptr() if rt is not None else None, n,
                                                count.data_ptr(), hist.data_ptr() if hist is not None else None,
                                                sums.data_ptr() if sums is not None else None,
                                                rows.data_ptr() if rows is not None else None, max_rows, self._s()),
                  "synseg_hsv_mask_hist")
        return dict(count=count, hist=hist, chan_sum=sums, row_count=rows)

    def hsv_mask_gather(self, rgb: torch.Tensor, roi, row_prefix: torch.Tensor, ranks: torch.Tensor) -> torch.Tensor:
        """RGB [n,3] of the ranks-th masked pixels (raster order) of one region."""
        n = ranks.numel()
        out = torch.empty((n, 3), dtype=torch.uint8, device=rgb.device)
        r = Roi(*[int(v) for v in roi]) if roi is not None else None
        if n:
            check(self.lib.synseg_hsv_mask_gather(self._h, C.byref(self._img(rgb, 3)), C.byref(r) if r is not None else None,
                                                  row_prefix.data_ptr(), ranks.data_ptr(), n, out.data_ptr(), self._s()),
                  "synseg_hsv_mask_gather")
        return out

    # ---- perceptual hash ------------------------------------------------------------------------
    def phash(self, src: torch.Tensor, src_kind: int = 0, rois=None) -> torch.Tensor:
        """int64 [n] holding the 64-bit hashes (reinterpret as uint64)."""
        ch = 3 if src_kind else 1
        s = _as3(src, ch)
        rt, n, a = self._rois_dev(rois, src.device)
        self._check_rois(a, s.shape[0], s.shape[1], s.shape[2])
        if rois is None:
            n = s.shape[0]
        out = torch.empty(n, dtype=torch.int64, device=src.device)
        if n:
            check(self.lib.synseg_phash(self._h, C.byref(self._img(src, ch)), src_kind, rt.data_ptr() if rt is not None else None, n,
                                        out.data_ptr(), self._s()), "synseg_phash")
        return out

    def select_rois(self, n_labels: torch.Tensor, stats: torch.Tensor, page_base: int, min_area: int, max_area: int, min_w: int,
                    min_h: int, rois: torch.Tensor, keys: torch.Tensor, count: torch.Tensor):
        """Device-side candidate selection: appends (image,x,y,w,h) rows to `rois` (int32 [cap,5]) and keys (int64 [cap])."""
        b, ml = stats.shape[0], stats.shape[1]
        check(self.lib.synseg_select_rois(self._h, n_labels.data_ptr(), stats.data_ptr(), b, ml, page_base, min_area, max_area, min_w, min_h,
                                          rois.data_ptr(), keys.data_ptr(), count.data_ptr(), rois.shape[0], self._s()), "synseg_select_rois")

    def phash_indirect(self, src: torch.Tensor, src_kind: int, rois: torch.Tensor, count: torch.Tensor, out: torch.Tensor):
        ch = 3 if src_kind else 1
        check(self.lib.synseg_phash_indirect(self._h, C.byref(self._img(src, ch)), src_kind, rois.data_ptr(), count.data_ptr(), rois.shape[0],
                                             out.data_ptr(), self._s()), "synseg_phash_indirect")

    def phash_dedup(self, hashes: torch.Tensor, keys: torch.Tensor, max_hamming: int = 4) -> torch.Tensor:
        n = hashes.numel()
        keep = torch.empty(n, dtype=torch.uint8, device=hashes.device)
        if n:
            check(self.lib.synseg_phash_dedup(self._h, hashes.data_ptr(), keys.data_ptr(), n, max_hamming, keep.data_ptr(), self._s()),
                  "synseg_phash_dedup")
        return keep

    # ---- fused pipelines ------------------------------------------------------------------------
    def _page_img(self, pages: torch.Tensor):
        """(synseg_img, channels, batch) of RGB pages [B,H,W,3] / [H,W,3] or grey pages [B,H,W] (channels from the layout)."""
        if pages.dim() == 4 or (pages.dim() == 3 and pages.shape[-1] == 3 and pages.stride(-1) == 1 and pages.stride(-2) == 3):
            r = _as3(pages, 3)
            return self._img(pages, 3), 3, r.shape[0]
        r = _as3(pages, 1)
        return self._img(pages, 1), 1, r.shape[0]

    def detect_pages(self, rgb: torch.Tensor, block_size: int, c: int, k: int, canny_lo: int = 50, canny_hi: int = 150,
                     max_labels: int = 1024, gray_out: Optional[torch.Tensor] = None, out=None, channels: Optional[int] = None):
        """Pages -> (n_labels int32 [B], stats int32 [B,max,5], centroids f64 [B,max,2]).
        RGB pages [B,H,W,3] (channels 3) or grey 'L' pages [B,H,W] (channels 1; guessed from the layout when None)."""
        if channels is None:
            img, channels, b = self._page_img(rgb)
        else:
            img, b = self._img(rgb, channels), _as3(rgb, channels).shape[0]
        dev = rgb.device
        if out is None:
            n = torch.empty(b, dtype=torch.int32, device=dev)
            stats = torch.empty((b, max_labels, 5), dtype=torch.int32, device=dev)
            cent = torch.empty((b, max_labels, 2), dtype=torch.float64, device=dev)
        else:
            n, stats, cent = out
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels, channels, 0)
        gimg = C.byref(self._img(gray_out)) if gray_out is not None else None
        check(self.lib.synseg_detect_pages(self._h, C.byref(img), C.byref(prm), gimg, n.data_ptr(), stats.data_ptr(),
                                           cent.data_ptr(), self._s()), "synseg_detect_pages")
        return n, stats, cent

    @staticmethod
    def region_buffers(batch: int, max_labels: int, max_regions: int, device=None, pinned: bool = False):
        """Result tensors of detect_regions / detect_regions_host: dict(n_labels, stats, regions (uint8 [B,R,72]), n_regions, flags)."""
        kw = dict(pin_memory=True) if pinned else dict(device=device)
        return dict(n_labels=torch.empty(batch, dtype=torch.int32, **kw), stats=torch.empty((batch, max_labels, 5), dtype=torch.int32, **kw),
                    regions=torch.empty((batch, max_regions, _lib.REGION_BYTES), dtype=torch.uint8, **kw),
                    n_regions=torch.empty(batch, dtype=torch.int32, **kw), flags=torch.empty(batch, dtype=torch.int32, **kw))

    @staticmethod
    def regions_view(regions: torch.Tensor) -> np.ndarray:
        """Host uint8 [B,R,72] tensor -> structured numpy view [B,R] (REGION_DTYPE), no copy."""
        a = regions.numpy()
        return a.view(REGION_DTYPE).reshape(a.shape[0], a.shape[1])

    def detect_regions(self, pages: torch.Tensor, block_size: int, c: int, k: int, dpi: float, page_width_pt: float, page_height_pt: float,
                       canny_lo: int = 50, canny_hi: int = 150, max_labels: int = 1024, max_regions: int = 64, min_extent_pt: float = 50.0,
                       out=None, channels: Optional[int] = None):
        """Device pages -> component tables + candidate regions with crop moments, one call on the current stream
        (synseg_detect_regions).  Returns the dict of `region_buffers` (device tensors)."""
        if channels is None:
            img, channels, b = self._page_img(pages)
        else:
            img, b = self._img(pages, channels), _as3(pages, channels).shape[0]
        o = out or self.region_buffers(b, max_labels, max_regions, pages.device)
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels, channels, 0)
        rp = RegionParams(float(dpi), float(page_width_pt), float(page_height_pt), float(min_extent_pt), max_regions, 0)
        check(self.lib.synseg_detect_regions(self._h, C.byref(img), C.byref(prm), C.byref(rp), o["n_labels"].data_ptr(), o["stats"].data_ptr(), None,
                                             o["regions"].data_ptr(), o["n_regions"].data_ptr(), o["flags"].data_ptr(), self._s()),
              "synseg_detect_regions")
        return o

    def regions_from_stats(self, n_labels: torch.Tensor, stats: torch.Tensor, pages: torch.Tensor, dpi: float, page_width_pt: float,
                           page_height_pt: float, max_regions: int = 64, min_extent_pt: float = 50.0, channels: Optional[int] = None):
        """Component tables -> (regions uint8 [B,R,72], n_regions, flags) on the device (synseg_regions_from_stats)."""
        if channels is None:
            img, channels, b = self._page_img(pages)
        else:
            img, b = self._img(pages, channels), _as3(pages, channels).shape[0]
        dev = pages.device
        regions = torch.empty((b, max_regions, _lib.REGION_BYTES), dtype=torch.uint8, device=dev)
        n_regions = torch.empty(b, dtype=torch.int32, device=dev)
        flags = torch.empty(b, dtype=torch.int32, device=dev)
        rp = RegionParams(float(dpi), float(page_width_pt), float(page_height_pt), float(min_extent_pt), max_regions, 0)
        check(self.lib.synseg_regions_from_stats(self._h, n_labels.data_ptr(), stats.data_ptr(), stats.shape[1], C.byref(img), channels, C.byref(rp),
                                                 regions.data_ptr(), n_regions.data_ptr(), flags.data_ptr(), self._s()), "synseg_regions_from_stats")
        return regions, n_regions, flags

    @staticmethod
    def _host_pages_layout(pages: torch.Tensor):
        if pages.is_cuda or pages.dtype != torch.uint8 or pages.stride(-1) != 1:
            raise ValueError("pages must be a CPU uint8 tensor [N,H,W,3] (RGB) or [N,H,W] (grey) with contiguous pixels")
        if pages.dim() == 4:
            if pages.shape[-1] != 3 or pages.stride(-2) != 3:
                raise ValueError("RGB pages must be interleaved [N,H,W,3]")
            ch = 3
        elif pages.dim() == 3:
            ch = 1
        else:
            raise ValueError("pages must be [N,H,W,3] or [N,H,W]")
        n, h, w = pages.shape[0], pages.shape[1], pages.shape[2]
        return ch, n, h, w, pages.stride(1), (pages.stride(0) if n > 1 else pages.stride(1) * h)

    def detect_pages_host(self, pages: torch.Tensor, block_size: int, c: int, k: int, canny_lo: int = 50, canny_hi: int = 150,
                          max_labels: int = 1024, chunk_pages: int = 16, want_centroids: bool = True, out=None):
        """HOST pages (CPU u8 tensor [N,H,W,3] RGB or [N,H,W] grey, ideally pinned) -> HOST tables (pinned): (n_labels [N], stats [N,max,5],
        centroids [N,max,2] | None).  Staging, H2D/compute/D2H overlap and chunking happen inside the library
        (synseg_detect_pages_host); the call is asynchronous on the current stream -- synchronise before reading."""
        ch, n, h, w, rs, ps = self._host_pages_layout(pages)
        if out is not None:                       # caller-provided (pinned) result tensors, reused across calls
            n_labels, stats, cent = out
        else:
            n_labels = torch.empty(n, dtype=torch.int32).pin_memory()
            stats = torch.empty((n, max_labels, 5), dtype=torch.int32).pin_memory()
            cent = torch.empty((n, max_labels, 2), dtype=torch.float64).pin_memory() if want_centroids else None
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels, ch, 0)
        check(self.lib.synseg_detect_pages_host(self._h, C.c_void_p(pages.data_ptr()), w, h, rs, ps,
                                                n, C.byref(prm), chunk_pages, C.c_void_p(n_labels.data_ptr()), C.c_void_p(stats.data_ptr()),
                                                C.c_void_p(cent.data_ptr()) if cent is not None else None, self._s()), "synseg_detect_pages_host")
        return n_labels, stats, cent

    def detect_regions_host(self, pages: torch.Tensor, block_size: int, c: int, k: int, dpi: float, page_width_pt: float, page_height_pt: float,
                            canny_lo: int = 50, canny_hi: int = 150, max_labels: int = 1024, max_regions: int = 64, min_extent_pt: float = 50.0,
                            chunk_pages: int = 10, out=None, want_stats: bool = True):
        """HOST pages (pinned u8 [N,H,W,3] RGB or [N,H,W] grey) -> HOST result tensors (dict of `region_buffers(pinned=True)`):
        staging ring, copy stream, detection, region rules and crop moments all inside the library (synseg_detect_regions_host).
        Asynchronous on the current stream -- synchronise (or record / wait an event) before reading."""
        ch, n, h, w, rs, ps = self._host_pages_layout(pages)
        o = out or self.region_buffers(n, max_labels, max_regions, pinned=True)
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels, ch, 0)
        rp = RegionParams(float(dpi), float(page_width_pt), float(page_height_pt), float(min_extent_pt), max_regions, 0)
        check(self.lib.synseg_detect_regions_host(self._h, C.c_void_p(pages.data_ptr()), w, h, rs, ps, n, C.byref(prm), C.byref(rp), chunk_pages,
                                                  C.c_void_p(o["n_labels"].data_ptr()), C.c_void_p(o["stats"].data_ptr()) if want_stats else None,
                                                  C.c_void_p(o["regions"].data_ptr()), C.c_void_p(o["n_regions"].data_ptr()),
                                                  C.c_void_p(o["flags"].data_ptr()), self._s()), "synseg_detect_regions_host")
        return o

    def hints_crops(self, packed: torch.Tensor, crops, kw: int = 25, kh: int = 25) -> torch.Tensor:
        """Ragged batch of crops packed in one CUDA u8 buffer.  crops = [(offset, width, height, row_stride, channels), ...].
        Returns int64 [n, 8] = h_count, v_count, edge_px, sum, sum_sq, non_zero, mask_px, 0 (see synseg_hints_crops)."""
        n = len(crops)
        out = torch.empty((n, 8), dtype=torch.int64, device=packed.device)
        if n == 0:
            return out
        if not packed.is_cuda or packed.dtype != torch.uint8 or not packed.is_contiguous():
            raise ValueError("packed must be a contiguous CUDA uint8 tensor")
        arr = (Crop * n)()
        total = packed.numel()
        for i, (off, w, h, rs, ch) in enumerate(crops):
            if off < 0 or w <= 0 or h <= 0 or rs < w * ch or off + rs * (h - 1) + w * ch > total:
                raise ValueError(f"crop {i} lies outside the packed buffer")
            arr[i] = Crop(int(off), int(w), int(h), int(rs), int(ch), 0)
        check(self.lib.synseg_hints_crops(self._h, C.c_void_p(packed.data_ptr()), arr, n, kw, kh, C.c_void_p(out.data_ptr()), self._s()),
              "synseg_hints_crops")
        return out

    def _crop_array(self, packed: torch.Tensor, crops):
        if not packed.is_cuda or packed.dtype != torch.uint8 or not packed.is_contiguous():
            raise ValueError("packed must be a contiguous CUDA uint8 tensor")
        n = len(crops)
        arr = (Crop * n)()
        total = packed.numel()
        for i, (off, w, h, rs, ch) in enumerate(crops):
            if off < 0 or w <= 0 or h <= 0 or rs < w * ch or off + rs * (h - 1) + w * ch > total:
                raise ValueError(f"crop {i} lies outside the packed buffer")
            arr[i] = Crop(int(off), int(w), int(h), int(rs), int(ch), 0)
        return arr

    def colors_crops(self, packed: torch.Tensor, crops, n_colors: int = 5, iters: int = 20, min_pixels: int = 100,
                     want_hist: bool = False):
        """Dominant colours of a ragged batch of crops (see synseg_colors_crops).  crops as for `hints_crops`.
        Returns (int64 [n, 2 + n_colors] = mask_px, k, (cluster pixels << 24 | R << 16 | G << 8 | B) x k; hist int32 [n, 4096] | None)."""
        n = len(crops)
        out = torch.empty((n, 2 + n_colors), dtype=torch.int64, device=packed.device)
        hist = torch.empty((n, HIST_BINS), dtype=torch.int32, device=packed.device) if want_hist else None
        if n == 0:
            return out, hist
        arr = self._crop_array(packed, crops)
        check(self.lib.synseg_colors_crops(self._h, C.c_void_p(packed.data_ptr()), arr, n, n_colors, iters, min_pixels,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(hist.data_ptr()) if hist is not None else None, self._s()),
              "synseg_colors_crops")
        return out, hist

    def hints_rois(self, pages: torch.Tensor, rois, kw: int = 25, kh: int = 25, channels: Optional[int] = None) -> torch.Tensor:
        """`hints_crops` on regions [(page, x, y, w, h), ...] of pages already on the device (read in place).  int64 [n, 8]."""
        if channels is None:
            img, channels, b = self._page_img(pages)
        else:
            img, b = self._img(pages, channels), _as3(pages, channels).shape[0]
        a = rois_array(rois)
        n = a.shape[0]
        out = torch.empty((n, 8), dtype=torch.int64, device=pages.device)
        if n:
            p = _as3(pages, channels)
            self._check_rois(a, p.shape[0], p.shape[1], p.shape[2])
            check(self.lib.synseg_hints_rois(self._h, C.byref(img), channels, a.ctypes.data_as(C.POINTER(Roi)), n, kw, kh, C.c_void_p(out.data_ptr()),
                                             self._s()), "synseg_hints_rois")
        return out

    def colors_rois(self, pages: torch.Tensor, rois, n_colors: int = 5, iters: int = 20, min_pixels: int = 100) -> torch.Tensor:
        """`colors_crops` on regions of RGB pages already on the device.  int64 [n, 2 + n_colors]."""
        a = rois_array(rois)
        n = a.shape[0]
        out = torch.empty((n, 2 + n_colors), dtype=torch.int64, device=pages.device)
        if n:
            p = _as3(pages, 3)
            self._check_rois(a, p.shape[0], p.shape[1], p.shape[2])
            check(self.lib.synseg_colors_rois(self._h, C.byref(self._img(pages, 3)), a.ctypes.data_as(C.POINTER(Roi)), n, n_colors, iters, min_pixels,
                                              C.c_void_p(out.data_ptr()), None, self._s()), "synseg_colors_rois")
        return out

    def grid_counts(self, src: torch.Tensor, rois=None, gray_mode: int = GRAY_PIL, kw: int = 25, kh: int = 25,
                    want_edges: bool = False, channels: Optional[int] = None):
        """Per region (h_count, v_count, edge_px) int64 [n,3] (+ edges u8 [n,maxH,maxW] when want_edges)."""
        ch = channels or (3 if (src.dim() >= 3 and src.shape[-1] == 3 and src.stride(-1) == 1 and src.stride(-2) == 3) else 1)
        s = _as3(src, ch)
        if rois is None:
            rois = [(i, 0, 0, s.shape[2], s.shape[1]) for i in range(s.shape[0])]
        a = rois_array(rois)
        self._check_rois(a, s.shape[0], s.shape[1], s.shape[2])
        n = a.shape[0]
        out = torch.empty((n, 3), dtype=torch.int64, device=src.device)
        edges = None
        if n == 0:
            return out, edges
        eimg = None
        if want_edges:
            edges = empty_plane(n, int(a[:, 4].max()), int(a[:, 3].max()), src.device)
            edges.zero_()
            eimg = C.byref(self._img(edges))
        check(self.lib.synseg_grid_counts(self._h, C.byref(self._img(src, ch)), ch, gray_mode, a.ctypes.data_as(C.POINTER(Roi)), n, kw, kh,
                                          out.data_ptr(), eimg, self._s()), "synseg_grid_counts")
        return out, edges


class PageSlots:
    """Renderer-facing pinned page slots (synseg_page_slot_*): a rasteriser writes pages straight into pinned memory owned
    by the library -- the handoff of pdf_image_segmentation.py:3638-3657 without the PNG round trip.

        slots = PageSlots(ctx, width, height, channels=1, pages_per_slot=16)
        s, pages = slots.acquire()            # numpy uint8 view [pages_per_slot, H, W(, 3)] of the pinned slot
        pages[:n] = ...                       # the renderer fills it
        slots.submit(s, n, block_size, C, k, dpi, page_w_pt, page_h_pt)
        res = slots.wait(s)                   # dict of numpy views: n_labels, stats, regions (REGION_DTYPE), n_regions, flags
    """

    def __init__(self, ctx: Context, width: int, height: int, channels: int = 3, pages_per_slot: int = 16, n_slots: int = 3,
                 max_labels: int = 1024, max_regions: int = 64):
        self.ctx, self.w, self.h, self.ch, self.pages, self.n_slots = ctx, width, height, channels, pages_per_slot, n_slots
        self.max_labels, self.max_regions = max_labels, max_regions
        check(ctx.lib.synseg_page_slots_init(ctx._h, width, height, channels, pages_per_slot, n_slots, max_labels, max_regions), "synseg_page_slots_init")
        self._n = {}

    @property
    def numa_node(self) -> int:
        return int(self.ctx.lib.synseg_page_slots_numa_node(self.ctx._h))

    def acquire(self):
        slot, ptr, rs, ps = C.c_int32(), C.c_void_p(), C.c_int64(), C.c_int64()
        check(self.ctx.lib.synseg_page_slot_acquire(self.ctx._h, C.byref(slot), C.byref(ptr), C.byref(rs), C.byref(ps)), "synseg_page_slot_acquire")
        buf = (C.c_uint8 * (ps.value * self.pages)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(self.pages, self.h, rs.value)[:, :, :self.w * self.ch]
        return slot.value, (a.reshape(self.pages, self.h, self.w, 3) if self.ch == 3 else a)

    def submit(self, slot: int, n_pages: int, block_size: int, c: int, k: int, dpi: float, page_width_pt: float, page_height_pt: float,
               canny_lo: int = 50, canny_hi: int = 150, min_extent_pt: float = 50.0):
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, self.max_labels, self.ch, 0)
        rp = RegionParams(float(dpi), float(page_width_pt), float(page_height_pt), float(min_extent_pt), self.max_regions, 0)
        check(self.ctx.lib.synseg_page_slot_submit(self.ctx._h, slot, n_pages, C.byref(prm), C.byref(rp), self.ctx._s()), "synseg_page_slot_submit")
        self._n[slot] = n_pages

    def wait(self, slot: int):
        p = [C.c_void_p() for _ in range(5)]
        check(self.ctx.lib.synseg_page_slot_wait(self.ctx._h, slot, *[C.byref(q) for q in p]), "synseg_page_slot_wait")
        n = self._n.pop(slot)

        def view(ptr, dtype, shape):
            count = int(np.prod(shape))
            buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr.value)
            return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        return dict(n_regions=view(p[0], np.int32, (n,)), flags=view(p[1], np.int32, (n,)), regions=view(p[2], REGION_DTYPE, (n, self.max_regions)),
                    n_labels=view(p[3], np.int32, (n,)), stats=view(p[4], np.int32, (n, self.max_labels, 5)))

    def close(self):
        if self.ctx is not None and getattr(self.ctx, "_h", None):
            self.ctx.lib.synseg_page_slots_release(self.ctx._h)
        self.ctx = None
