"""ctypes binding of libsynseg.so (include/synseg.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SYNSEG_LIB") or os.path.join(_HERE, "libsynseg.so")   # SYNSEG_LIB: a variant built with SYNSEG_BUILD_TAG (tuning)


class Img(C.Structure):
    """synseg_img"""
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_stride", C.c_int64),
                ("batch", C.c_int32), ("_pad", C.c_int32), ("batch_stride", C.c_int64)]


class Roi(C.Structure):
    """synseg_roi"""
    _fields_ = [("image", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class Crop(C.Structure):
    """synseg_crop"""
    _fields_ = [("offset", C.c_uint64), ("width", C.c_int32), ("height", C.c_int32), ("row_stride", C.c_int64),
                ("channels", C.c_int32), ("_pad", C.c_int32)]


class DetectParams(C.Structure):
    """synseg_detect_params"""
    _fields_ = [("block_size", C.c_int32), ("C", C.c_int32), ("canny_lo", C.c_int32), ("canny_hi", C.c_int32),
                ("k", C.c_int32), ("max_labels", C.c_int32), ("channels", C.c_int32), ("_pad", C.c_int32)]


class Region(C.Structure):
    """synseg_region (72 bytes)"""
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("x1", C.c_double), ("y1", C.c_double),
                ("px", C.c_int32), ("py", C.c_int32), ("pw", C.c_int32), ("ph", C.c_int32),
                ("kind", C.c_int32), ("count", C.c_int32), ("sum", C.c_uint64), ("sum_sq", C.c_uint64)]


class RegionParams(C.Structure):
    """synseg_region_params"""
    _fields_ = [("dpi", C.c_double), ("page_width_pt", C.c_double), ("page_height_pt", C.c_double), ("min_extent_pt", C.c_double),
                ("max_regions", C.c_int32), ("_pad", C.c_int32)]


REGION_BYTES = 72
REGION_CC, REGION_CLUSTER = 1, 2
FLAG_LABELS, FLAG_AMBIGUOUS, FLAG_CAPACITY = 1, 2, 4


# name -> (restype, argtypes); mirrors include/synseg.h declaration by declaration
_P = C.POINTER
SIGNATURES = {
    "synseg_version": (C.c_int, []),
    "synseg_last_error": (C.c_char_p, []),
    "synseg_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "synseg_destroy": (C.c_int, [C.c_void_p]),
    "synseg_reserve": (C.c_int, [C.c_void_p, C.c_size_t]),
    "synseg_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "synseg_launch_count": (C.c_int64, [C.c_void_p]),
    "synseg_guard_violations": (C.c_int64, [C.c_void_p, _P(C.c_int64), _P(C.c_int32)]),
    "synseg_profile_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "synseg_profile_end": (C.c_int, [C.c_void_p, _P(C.c_char_p), _P(C.c_float), C.c_int]),
    "synseg_select_rois": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "synseg_phash_indirect": (C.c_int, [C.c_void_p, _P(Img), C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "synseg_rgb2gray": (C.c_int, [C.c_void_p, _P(Img), _P(Img), C.c_int, C.c_void_p]),
    "synseg_adaptive_mean": (C.c_int, [C.c_void_p, _P(Img), _P(Img), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "synseg_canny": (C.c_int, [C.c_void_p, _P(Img), _P(Img), C.c_int, C.c_int, C.c_void_p]),
    "synseg_morph": (C.c_int, [C.c_void_p, _P(Img), _P(Img), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "synseg_ccl_stats": (C.c_int, [C.c_void_p, _P(Img), _P(Img), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "synseg_moments": (C.c_int, [C.c_void_p, _P(Img), C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "synseg_hsv_mask_hist": (C.c_int, [C.c_void_p, _P(Img), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "synseg_hsv_mask_gather": (C.c_int, [C.c_void_p, _P(Img), _P(Roi), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "synseg_phash": (C.c_int, [C.c_void_p, _P(Img), C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "synseg_phash_dedup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "synseg_detect_pages": (C.c_int, [C.c_void_p, _P(Img), _P(DetectParams), _P(Img), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_detect_pages_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P(DetectParams), C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_regions_from_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _P(Img), C.c_int, _P(RegionParams), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "synseg_detect_regions": (C.c_int, [C.c_void_p, _P(Img), _P(DetectParams), _P(RegionParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_detect_regions_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P(DetectParams),
                                             _P(RegionParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_page_slots_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "synseg_page_slot_acquire": (C.c_int, [C.c_void_p, _P(C.c_int32), _P(C.c_void_p), _P(C.c_int64), _P(C.c_int64)]),
    "synseg_page_slot_submit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _P(DetectParams), _P(RegionParams), C.c_void_p]),
    "synseg_page_slot_wait": (C.c_int, [C.c_void_p, C.c_int32, _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p)]),
    "synseg_page_slots_release": (C.c_int, [C.c_void_p]),
    "synseg_page_slots_numa_node": (C.c_int, [C.c_void_p]),
    "synseg_hints_crops": (C.c_int, [C.c_void_p, C.c_void_p, _P(Crop), C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "synseg_colors_crops": (C.c_int, [C.c_void_p, C.c_void_p, _P(Crop), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_comm_unique_id": (C.c_int, [C.c_void_p]),
    "synseg_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "synseg_comm_destroy": (C.c_int, [C.c_void_p]),
    "synseg_comm_info": (C.c_int, [C.c_void_p, _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]),
    "synseg_dedup_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "synseg_hints_rois": (C.c_int, [C.c_void_p, _P(Img), C.c_int, _P(Roi), C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "synseg_colors_rois": (C.c_int, [C.c_void_p, _P(Img), _P(Roi), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "synseg_grid_counts": (C.c_int, [C.c_void_p, _P(Img), C.c_int, C.c_int, _P(Roi), C.c_int32, C.c_int, C.c_int, C.c_void_p, _P(Img), C.c_void_p]),
}

_lib = None


class SynsegError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libsynseg.so.  Raises if it has not been built: the CUDA library IS the product."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m synapta_image_segmentation_b200.build` "
                "(there is no CPU fallback for the region-detection hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().synseg_last_error().decode("utf-8", "replace")
        raise SynsegError(f"{what or 'libsynseg'} failed (code {rc}): {msg}")
