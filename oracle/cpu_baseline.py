"""CPU timing of the reference's OpenCV path -- used ONLY by bench.py's cpu_baseline / --impl reference legs.

The reference has no raster detector of its own; its CPU path for this stage is the cv2 4.13.0 primitive chain
the GPU kernels are bit-exact against (oracle/cv2_chain.py: SURVEY.md 8d).  Two arrangements are timed and the
faster one is reported: (i) one page at a time with cv2's own thread pool on all cores; (ii) one page per
Python thread with cv2 single-threaded (cv2 releases the GIL inside every primitive) -- BASELINE.md section 2.
No fork/multiprocessing: forking a process that has initialised cv2 / torch thread pools can deadlock.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def _chain_all_threads(pages: np.ndarray, dpi: int) -> float:
    import cv2
    from oracle import cv2_chain
    cv2.setNumThreads(os.cpu_count() or 1)
    cv2_chain.page_chain(pages[0], dpi)        # warm-up
    t0 = time.perf_counter()
    for p in pages:
        cv2_chain.page_chain(p, dpi)
    return (time.perf_counter() - t0) / len(pages)


def _chain_page_parallel(pages: np.ndarray, dpi: int, workers: int) -> float:
    import cv2
    from oracle import cv2_chain
    cv2.setNumThreads(1)
    try:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda p: cv2_chain.page_chain(p, dpi)["n"], pages[:workers]))   # warm-up
            t0 = time.perf_counter()
            list(ex.map(lambda p: cv2_chain.page_chain(p, dpi)["n"], pages))
            return (time.perf_counter() - t0) / len(pages)
    finally:
        cv2.setNumThreads(os.cpu_count() or 1)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def measure_pages(pages: np.ndarray, dpi: int):
    """pages: u8 [n,H,W,3].  Returns dict(value pages/s, cores, arrangement, pages, other)."""
    cores = os.cpu_count() or 1
    a = _chain_all_threads(pages[:max(2, min(len(pages), 8))], dpi)
    b = _chain_page_parallel(pages, dpi, cores)
    if b < a:
        return dict(value=1.0 / b, cores=cores, arrangement="one page per thread, cv2 single-threaded", pages=len(pages), other=1.0 / a)
    return dict(value=1.0 / a, cores=cores, arrangement="one page at a time, cv2 thread pool on all cores", pages=len(pages), other=1.0 / b)
