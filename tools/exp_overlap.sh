#!/bin/bash
# overlap experiment: resident bench for (SYNSEG_OVERLAP chunks, SYNSEG_STREAMS) combinations
for cfg in "2 2" "3 3" "4 4" "4 2" "6 3"; do
  set -- $cfg
  echo "== SYNSEG_OVERLAP=$1 SYNSEG_STREAMS=$2"
  if [ "$cfg" = "3 3" ]; then SYNSEG_OVERLAP=$1 SYNSEG_STREAMS=$2 timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2; fi
  SYNSEG_OVERLAP=$1 SYNSEG_STREAMS=$2 timeout 200 python bench.py --no-cpu --no-e2e 2>gpurun_out/exp_overlap_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'value', d['value'])"
done
