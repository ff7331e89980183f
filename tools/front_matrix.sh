#!/bin/bash
# step time of the resident pipeline for the fused / separate variants of the front end and the morphology, with one chain and two chains
cd "$(dirname "$0")/.."
for tma in 0 1; do for fm in 0 1; do for ov in 1 2; do
  env_tma=""; [ $tma = 0 ] && env_tma="SYNSEG_NO_TMA=1"
  env_fm=""; [ $fm = 0 ] && env_fm="SYNSEG_NO_FUSED_MORPH=1"
  env $env_tma $env_fm SYNSEG_OVERLAP=$ov python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('tma $tma fused_morph $fm chains $ov: step', round(d['ms_per_step'],4), 'pages/s', round(d['value']), 'serial kernels', round(d['roofline']['serial_step_ms'],4))"
done; done; done
