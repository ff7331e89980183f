#!/bin/sh
# Experiment driver (run on the GPU box): warps per CTA x band heights for the two front-end stencil kernels.
# Results of the round-1 sweep: DESIGN.md section 6 (1-warp CTAs, Canny bands of 32 rows, threshold bands of 64 rows).
for w in 1 2 4; do
 SYNSEG_NVCC_EXTRA="-DSYNSEG_CN_WARPS=$w -DSYNSEG_AD_WARPS=$w" python -m synapta_image_segmentation_b200.build --force > /dev/null 2>&1
 for cfg in "16 48" "32 64" "64 128" "128 320"; do
  set -- $cfg
  SYNSEG_TUNE_CANNY_BAND=$1 SYNSEG_TUNE_AD_BAND=$2 python bench.py --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['roofline']['kernels'];print('warps',$w,'canny band',$1,'ad band',$2,'step',round(d['ms_per_step'],3),'canny',k['canny_classes']['ms_per_step'],'adaptive',k['adaptive_mean']['ms_per_step'])"
 done
done
python -m synapta_image_segmentation_b200.build --force > /dev/null 2>&1
