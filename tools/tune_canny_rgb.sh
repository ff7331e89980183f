#!/bin/bash
# Rebuilds canny.cu with different TMA ring depths on the GPU box and times the kernel for several band heights (ms per 50 pages).
cd "$(dirname "$0")/.."
for v in "2 0" "3 0"; do
  set -- $v
  touch synapta_image_segmentation_b200/csrc/canny.cu
  SYNSEG_NVCC_EXTRA="-DSYNSEG_CR_DEPTH=$1 -DSYNSEG_CR_PREFETCH=$2" python -m synapta_image_segmentation_b200.build > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  for band in 24 32 48 64 96; do
  SYNSEG_TUNE_CANNY_BAND=$band python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('depth $1 band $band: canny_rgb', k['canny_rgb']['ms_per_step'], 'step', round(d['ms_per_step'],4), 'serial', round(d['roofline']['serial_step_ms'],4))"
  done
done
SYNSEG_NO_TMA=1 python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('separate kernels: rgb2gray', k['rgb2gray']['ms_per_step'], 'canny', k['canny_classes']['ms_per_step'], 'step', round(d['ms_per_step'],4), 'serial', round(d['roofline']['serial_step_ms'],4))"
touch synapta_image_segmentation_b200/csrc/canny.cu
python -m synapta_image_segmentation_b200.build > /dev/null 2>&1
