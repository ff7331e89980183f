"""CPU oracle for the region-detection hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``synapta_image_segmentation_b200``) never does.

Three layers, all CPU:

* ``synseg_oracle.c``  (this module's ctypes wrappers) -- plain-C restatement of every primitive;
* ``oracle.cv2_chain``  -- the same primitives through the live cv2 / PIL wheels (the libraries the
  reference itself calls: opencv-python 4.13.0, Pillow 12.2.0, numpy 2.3.5; un-vendored, unpinned
  by the reference), composed into the canonical page chain of SURVEY.md 8(d);
* ``oracle.ref_import`` -- imports ``/root/reference/pdf_image_segmentation.py`` with ``fitz`` and
  ``paddleocr`` stubbed (build container only) to generate golden vectors under ``tests/golden``.

Pinning status: the reference ships no tests or golden vectors for this path ("parity unpinned by
the reference's own tests").  The oracle is instead pinned against (a) the live
cv2/PIL primitives on the GPU box and here, (b) outputs of the imported reference functions,
committed as fixtures in ``tests/golden`` by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsynseg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synseg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i32p, u32p, u64p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_uint32, C.c_uint64, C.c_double))
        L.orc_rgb2gray_cv.argtypes = [u8p, C.c_int64, u8p]
        L.orc_rgb2gray_pil.argtypes = [u8p, C.c_int64, u8p]
        L.orc_hsv_sv.argtypes = [u8p, C.c_int64, u8p, u8p]
        L.orc_hsv_mask.argtypes = [u8p, C.c_int64, u8p]
        L.orc_hsv_mask.restype = C.c_int64
        L.orc_moments_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int64, u64p, u64p, u64p]
        L.orc_adaptive_mean.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_canny_classes.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_canny.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_morph_rect.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_morphology_ex.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_ccl8_stats.argtypes = [u8p, C.c_int, C.c_int, i32p, i32p, f64p, C.c_int32]
        L.orc_ccl8_stats.restype = C.c_int32
        L.orc_hsv_hist.argtypes = [u8p, C.c_int, C.c_int, C.c_int64, C.c_int, u32p, u64p]
        L.orc_hsv_hist.restype = C.c_int64
        L.orc_phash_basis.argtypes = [i32p]
        L.orc_phash.argtypes = [u8p, C.c_int, C.c_int, C.c_int64]
        L.orc_phash.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def rgb2gray_cv(rgb) -> np.ndarray:
    rgb = _u8(rgb); out = np.empty(rgb.shape[:-1], np.uint8)
    lib().orc_rgb2gray_cv(_p(rgb, C.c_uint8), out.size, _p(out, C.c_uint8)); return out


def rgb2gray_pil(rgb) -> np.ndarray:
    rgb = _u8(rgb); out = np.empty(rgb.shape[:-1], np.uint8)
    lib().orc_rgb2gray_pil(_p(rgb, C.c_uint8), out.size, _p(out, C.c_uint8)); return out


def hsv_sv(rgb):
    rgb = _u8(rgb); s = np.empty(rgb.shape[:-1], np.uint8); v = np.empty_like(s)
    lib().orc_hsv_sv(_p(rgb, C.c_uint8), s.size, _p(s, C.c_uint8), _p(v, C.c_uint8)); return s, v


def hsv_mask(rgb):
    rgb = _u8(rgb); m = np.empty(rgb.shape[:-1], np.uint8)
    n = lib().orc_hsv_mask(_p(rgb, C.c_uint8), m.size, _p(m, C.c_uint8)); return m, int(n)


def moments_u8(x):
    """(sum, sumsq, nonzero) as Python ints."""
    x = _u8(x); h, w = x.shape
    s, ss, nz = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().orc_moments_u8(_p(x, C.c_uint8), h, w, w, C.byref(s), C.byref(ss), C.byref(nz))
    return int(s.value), int(ss.value), int(nz.value)


def variance_from_moments(n: int, s: int, ss: int) -> float:
    """Exact population variance (n*ss - s*s)/n^2, correctly rounded (Python big-int division)."""
    return (n * ss - s * s) / (n * n)


def adaptive_mean(g, bs: int, c: int, inv: bool = True) -> np.ndarray:
    g = _u8(g); h, w = g.shape; out = np.empty_like(g)
    lib().orc_adaptive_mean(_p(g, C.c_uint8), h, w, bs, c, int(inv), _p(out, C.c_uint8)); return out


def canny_classes(g, lo=50, hi=150) -> np.ndarray:
    g = _u8(g); h, w = g.shape; out = np.empty_like(g)
    lib().orc_canny_classes(_p(g, C.c_uint8), h, w, lo, hi, _p(out, C.c_uint8)); return out


def canny(g, lo=50, hi=150) -> np.ndarray:
    g = _u8(g); h, w = g.shape; out = np.empty_like(g)
    lib().orc_canny(_p(g, C.c_uint8), h, w, lo, hi, _p(out, C.c_uint8)); return out


ERODE, DILATE, OPEN, CLOSE = 0, 1, 2, 3


def morph_rect(src, op: int, kw: int, kh: int, ax: int | None = None, ay: int | None = None) -> np.ndarray:
    src = _u8(src); h, w = src.shape; out = np.empty_like(src)
    ax = kw // 2 if ax is None else ax; ay = kh // 2 if ay is None else ay
    lib().orc_morph_rect(_p(src, C.c_uint8), h, w, op, kw, kh, ax, ay, _p(out, C.c_uint8)); return out


def morphology_ex(src, mop: int, kw: int, kh: int, iterations: int = 1) -> np.ndarray:
    src = _u8(src); h, w = src.shape; out = np.empty_like(src)
    lib().orc_morphology_ex(_p(src, C.c_uint8), h, w, mop, kw, kh, iterations, _p(out, C.c_uint8)); return out


def ccl8_stats(m, max_labels: int | None = None):
    """(n_labels, labels i32 [H,W], stats i32 [n,5], centroids f64 [n,2]) like cv2."""
    m = _u8(m); h, w = m.shape
    cap = max_labels or (h * w // 2 + 2)
    labels = np.empty((h, w), np.int32); stats = np.empty((cap, 5), np.int32); cent = np.empty((cap, 2), np.float64)
    n = lib().orc_ccl8_stats(_p(m, C.c_uint8), h, w, _p(labels, C.c_int32), _p(stats, C.c_int32), _p(cent, C.c_double), cap)
    if n < 0:
        raise RuntimeError("max_labels too small")
    return n, labels, stats[:n].copy(), cent[:n].copy()


def hsv_hist(rgb, bits: int = 5):
    rgb = _u8(rgb); h, w, _ = rgb.shape
    hist = np.zeros(1 << (3 * bits), np.uint32); sums = np.zeros((1 << (3 * bits), 3), np.uint64)
    n = lib().orc_hsv_hist(_p(rgb, C.c_uint8), h, w, 3 * w, bits, _p(hist, C.c_uint32), _p(sums, C.c_uint64))
    return int(n), hist, sums


def phash_basis() -> np.ndarray:
    b = np.empty((8, 32), np.int32); lib().orc_phash_basis(_p(b, C.c_int32)); return b


def phash(gray) -> int:
    gray = _u8(gray); h, w = gray.shape
    return int(lib().orc_phash(_p(gray, C.c_uint8), h, w, w))
