#!/bin/sh
# Experiment driver (GPU box): cost of the NMS loop / the per-pixel threshold test (results are wrong with the knobs on).
for f in 0 1 2; do
  SYNSEG_TUNE_FLAGS=$f python bench.py --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['roofline']['kernels'];print('flags',$f,'step',round(d['ms_per_step'],3),'canny',k['canny_classes']['ms_per_step'],'adaptive',k['adaptive_mean']['ms_per_step'])"
done
