#!/bin/bash
# Memory-safety run (compute-sanitizer is closed on this pool): builds libsynseg.so with -DSYNSEG_GUARD -- canary zones around every
# scratch allocation, compared whenever scratch is reused and when a public call returns; index assertions inside the kernels as real
# device asserts -- runs the whole GPU suite against it and restores the normal build.  Output: gpurun_out/guard_run.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SYNSEG_NVCC_EXTRA="-DSYNSEG_GUARD" python -m synapta_image_segmentation_b200.build --force > gpurun_out/guard_build.log 2>&1 || { echo "guard build failed"; tail -5 gpurun_out/guard_build.log; exit 1; }
python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^DEBUG" | tail -40 > gpurun_out/guard_run.log
grep -n "SYNSEG_GUARD\|passed\|failed" gpurun_out/guard_run.log
python -m synapta_image_segmentation_b200.build --force > /dev/null 2>&1
