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
from ._lib import Crop, DetectParams, Img, Roi, check

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

    # ---- per-kernel timing ---------------------------------------------------------------------
    def profile_begin(self):
        check(self.lib.synseg_profile_begin(self._h, _stream()), "synseg_profile_begin")

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
        check(self.lib.synseg_rgb2gray(self._h, C.byref(img_of(rgb, 3)), C.byref(img_of(out)), mode, _stream()), "synseg_rgb2gray")
        return out if rgb.dim() == 4 else out[0]

    # ---- threshold / edges ----------------------------------------------------------------------
    def adaptive_mean(self, gray: torch.Tensor, block_size: int, c: int, invert: bool = True,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = _as3(gray, 1)
        if out is None:
            out = empty_plane(*g.shape, gray.device)
        check(self.lib.synseg_adaptive_mean(self._h, C.byref(img_of(gray)), C.byref(img_of(out)), block_size, c, int(invert), _stream()),
              "synseg_adaptive_mean")
        return out if gray.dim() == 3 else out[0]

    def canny(self, gray: torch.Tensor, lo: int = 50, hi: int = 150, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = _as3(gray, 1)
        if out is None:
            out = empty_plane(*g.shape, gray.device)
        check(self.lib.synseg_canny(self._h, C.byref(img_of(gray)), C.byref(img_of(out)), lo, hi, _stream()), "synseg_canny")
        return out if gray.dim() == 3 else out[0]

    # ---- morphology ---------------------------------------------------------------------------
    def morph(self, src: torch.Tensor, op: int, kw: int, kh: int, anchor=(-1, -1), iterations: int = 1, binary: bool = False,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        s = _as3(src, 1)
        if out is None:
            out = empty_plane(*s.shape, src.device)
        check(self.lib.synseg_morph(self._h, C.byref(img_of(src)), C.byref(img_of(out)), op, kw, kh, anchor[0], anchor[1], iterations,
                                    1 if binary else 0, _stream()), "synseg_morph")
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
        limg = C.byref(img_of(labels, 1, 4)) if want_labels else None
        check(self.lib.synseg_ccl_stats(self._h, C.byref(img_of(mask)), limg, n.data_ptr(), stats.data_ptr(), cent.data_ptr(),
                                        max_labels, _stream()), "synseg_ccl_stats")
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
            check(self.lib.synseg_moments(self._h, C.byref(img_of(src, ch)), src_kind, rt.data_ptr() if rt is not None else None, n,
                                          out.data_ptr(), _stream()), "synseg_moments")
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
            check(self.lib.synseg_hsv_mask_hist(self._h, C.byref(img_of(rgb, 3)), rt.data_ptr() if rt is not None else None, n,
                                                count.data_ptr(), hist.data_ptr() if hist is not None else None,
                                                sums.data_ptr() if sums is not None else None,
                                                rows.data_ptr() if rows is not None else None, max_rows, _stream()),
                  "synseg_hsv_mask_hist")
        return dict(count=count, hist=hist, chan_sum=sums, row_count=rows)

    def hsv_mask_gather(self, rgb: torch.Tensor, roi, row_prefix: torch.Tensor, ranks: torch.Tensor) -> torch.Tensor:
        """RGB [n,3] of the ranks-th masked pixels (raster order) of one region."""
        n = ranks.numel()
        out = torch.empty((n, 3), dtype=torch.uint8, device=rgb.device)
        r = Roi(*[int(v) for v in roi]) if roi is not None else None
        if n:
            check(self.lib.synseg_hsv_mask_gather(self._h, C.byref(img_of(rgb, 3)), C.byref(r) if r is not None else None,
                                                  row_prefix.data_ptr(), ranks.data_ptr(), n, out.data_ptr(), _stream()),
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
            check(self.lib.synseg_phash(self._h, C.byref(img_of(src, ch)), src_kind, rt.data_ptr() if rt is not None else None, n,
                                        out.data_ptr(), _stream()), "synseg_phash")
        return out

    def select_rois(self, n_labels: torch.Tensor, stats: torch.Tensor, page_base: int, min_area: int, max_area: int, min_w: int,
                    min_h: int, rois: torch.Tensor, keys: torch.Tensor, count: torch.Tensor):
        """Device-side candidate selection: appends (image,x,y,w,h) rows to `rois` (int32 [cap,5]) and keys (int64 [cap])."""
        b, ml = stats.shape[0], stats.shape[1]
        check(self.lib.synseg_select_rois(self._h, n_labels.data_ptr(), stats.data_ptr(), b, ml, page_base, min_area, max_area, min_w, min_h,
                                          rois.data_ptr(), keys.data_ptr(), count.data_ptr(), rois.shape[0], _stream()), "synseg_select_rois")

    def phash_indirect(self, src: torch.Tensor, src_kind: int, rois: torch.Tensor, count: torch.Tensor, out: torch.Tensor):
        ch = 3 if src_kind else 1
        check(self.lib.synseg_phash_indirect(self._h, C.byref(img_of(src, ch)), src_kind, rois.data_ptr(), count.data_ptr(), rois.shape[0],
                                             out.data_ptr(), _stream()), "synseg_phash_indirect")

    def phash_dedup(self, hashes: torch.Tensor, keys: torch.Tensor, max_hamming: int = 4) -> torch.Tensor:
        n = hashes.numel()
        keep = torch.empty(n, dtype=torch.uint8, device=hashes.device)
        if n:
            check(self.lib.synseg_phash_dedup(self._h, hashes.data_ptr(), keys.data_ptr(), n, max_hamming, keep.data_ptr(), _stream()),
                  "synseg_phash_dedup")
        return keep

    # ---- fused pipelines ------------------------------------------------------------------------
    def detect_pages(self, rgb: torch.Tensor, block_size: int, c: int, k: int, canny_lo: int = 50, canny_hi: int = 150,
                     max_labels: int = 1024, gray_out: Optional[torch.Tensor] = None, out=None):
        """RGB pages [B,H,W,3] -> (n_labels int32 [B], stats int32 [B,max,5], centroids f64 [B,max,2])."""
        r = _as3(rgb, 3)
        b = r.shape[0]
        dev = rgb.device
        if out is None:
            n = torch.empty(b, dtype=torch.int32, device=dev)
            stats = torch.empty((b, max_labels, 5), dtype=torch.int32, device=dev)
            cent = torch.empty((b, max_labels, 2), dtype=torch.float64, device=dev)
        else:
            n, stats, cent = out
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels)
        gimg = C.byref(img_of(gray_out)) if gray_out is not None else None
        check(self.lib.synseg_detect_pages(self._h, C.byref(img_of(rgb, 3)), C.byref(prm), gimg, n.data_ptr(), stats.data_ptr(),
                                           cent.data_ptr(), _stream()), "synseg_detect_pages")
        return n, stats, cent

    def detect_pages_host(self, pages: torch.Tensor, block_size: int, c: int, k: int, canny_lo: int = 50, canny_hi: int = 150,
                          max_labels: int = 1024, chunk_pages: int = 16, want_centroids: bool = True, out=None):
        """HOST pages (CPU u8 tensor [N,H,W,3], ideally pinned) -> HOST tables (pinned): (n_labels [N], stats [N,max,5],
        centroids [N,max,2] | None).  Staging, H2D/compute/D2H overlap and chunking happen inside the library
        (synseg_detect_pages_host); the call is asynchronous on the current stream -- synchronise before reading."""
        if pages.is_cuda or pages.dtype != torch.uint8 or pages.dim() != 4 or pages.shape[-1] != 3 or pages.stride(-1) != 1 or pages.stride(-2) != 3:
            raise ValueError("pages must be a CPU uint8 tensor [N,H,W,3] with interleaved pixels")
        n, h, w, _ = pages.shape
        if out is not None:                       # caller-provided (pinned) result tensors, reused across calls
            n_labels, stats, cent = out
        else:
            n_labels = torch.empty(n, dtype=torch.int32).pin_memory()
            stats = torch.empty((n, max_labels, 5), dtype=torch.int32).pin_memory()
            cent = torch.empty((n, max_labels, 2), dtype=torch.float64).pin_memory() if want_centroids else None
        prm = DetectParams(block_size, c, canny_lo, canny_hi, k, max_labels)
        check(self.lib.synseg_detect_pages_host(self._h, C.c_void_p(pages.data_ptr()), w, h, pages.stride(1), pages.stride(0) if n > 1 else pages.stride(1) * h,
                                                n, C.byref(prm), chunk_pages, C.c_void_p(n_labels.data_ptr()), C.c_void_p(stats.data_ptr()),
                                                C.c_void_p(cent.data_ptr()) if cent is not None else None, _stream()), "synseg_detect_pages_host")
        return n_labels, stats, cent

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
        check(self.lib.synseg_hints_crops(self._h, C.c_void_p(packed.data_ptr()), arr, n, kw, kh, C.c_void_p(out.data_ptr()), _stream()),
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
                                           C.c_void_p(out.data_ptr()), C.c_void_p(hist.data_ptr()) if hist is not None else None, _stream()),
              "synseg_colors_crops")
        return out, hist

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
            eimg = C.byref(img_of(edges))
        check(self.lib.synseg_grid_counts(self._h, C.byref(img_of(src, ch)), ch, gray_mode, a.ctypes.data_as(C.POINTER(Roi)), n, kw, kh,
                                          out.data_ptr(), eimg, _stream()), "synseg_grid_counts")
        return out, edges
