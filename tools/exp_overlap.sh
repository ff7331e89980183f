#!/bin/bash
# overlap experiment: parity tests + resident bench for (SYNSEG_OVERLAP chunks, SYNSEG_FORK) combinations
for cfg in "1 0" "1 1" "2 0" "2 1" "3 1"; do
  set -- $cfg
  echo "== SYNSEG_OVERLAP=$1 SYNSEG_FORK=$2"
  if [ "$cfg" != "1 0" ] && [ "$cfg" != "2 0" ] && [ "$cfg" != "3 1" ]; then SYNSEG_OVERLAP=$1 SYNSEG_FORK=$2 timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2; fi
  SYNSEG_OVERLAP=$1 SYNSEG_FORK=$2 timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/exp_overlap_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'value', d['value'])"
done
