#!/bin/bash
# hysteresis sweeps: number of sweeps / rounds per band / band height against the step time (text pages and dense pages)
cd "$(dirname "$0")/.."
for v in "4 16 32" "3 16 32" "5 12 32" "4 24 32" "3 32 32" "4 12 16"; do
  set -- $v
  touch synapta_image_segmentation_b200/csrc/ccl.cu synapta_image_segmentation_b200/csrc/hyst_sweep.cu
  SYNSEG_NVCC_EXTRA="-DSYNSEG_HS_SWEEPS=$1 -DSYNSEG_HS_MAX_IT=$2 -DSYNSEG_HS_TR=$3" python -m synapta_image_segmentation_b200.build > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  python bench.py --no-cpu --no-corpus --crops 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
hy=sum(k[n]['ms_per_step'] for n in ('hyst_sweep','hyst_flag','hyst_final') if n in k)
print('sweeps $1 rounds $2 band $3: step', round(d['ms_per_step'],4), 'hyst_sweep', k['hyst_sweep']['ms_per_step'], 'hyst*', round(hy,4), 'merge', k['ccl_merge']['ms_per_step'], '| dense step', round(d['dense_pages']['ms_per_step'],4))"
done
touch synapta_image_segmentation_b200/csrc/ccl.cu synapta_image_segmentation_b200/csrc/hyst_sweep.cu
python -m synapta_image_segmentation_b200.build > /dev/null 2>&1
