"""Builds libsynseg.so (hand-written sm_100a CUDA kernels + the C ABI of include/synseg.h) in-tree with nvcc.

    python -m synapta_image_segmentation_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SYNSEG_BUILD_TAG=<tag> builds a variant (other -D flags through SYNSEG_NVCC_EXTRA) beside the product library:
# libsynseg_<tag>.so, selected at run time with SYNSEG_LIB=<path> (tuning experiments: the variants are built here and travel to the GPU box)
_TAG = os.environ.get("SYNSEG_BUILD_TAG", "")
OBJ = os.path.join(CSRC, "_obj" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libsynseg" + ("_" + _TAG if _TAG else "") + ".so")
SOURCES = ["ctx.cu", "gray.cu", "threshold.cu", "canny.cu", "morph.cu", "morph_fused.cu", "ccl.cu", "hyst_sweep.cu", "reduce.cu", "phash.cu", "colors.cu", "regions.cu", "exchange.cu", "pipeline.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"] + os.environ.get("SYNSEG_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "synseg.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), res.stderr))
                if verbose and res.stderr:
                    print(res.stderr, file=sys.stderr)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xlinker", "--exclude-libs,ALL", "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
