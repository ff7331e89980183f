#!/bin/bash
# session-3 experiment 1: compile-time-radius adaptive kernel + byte-2 grey packing; parity subset, A/B bench, full ncu capture with source
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py tests/test_gpu_properties.py -m gpu -q -x -k "adaptive or config1 or config2 or odd_sizes or edge_cases or variant or pipeline" 2>&1 | tail -2
B="python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e"
for v in "" "SYNSEG_AD_GENERIC=1"; do
  env $v $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('[$v] step', round(d['ms_per_step'],4), {n:k[n]['ms_per_step'] for n in ('canny_rgb','adaptive_mean')})"
done
SYNSEG_OVERLAP=1 $B --steps 2 --warmup 3 > gpurun_out/s3_plain.log 2>&1 &&
SYNSEG_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:'canny_rgb|adaptive_mean|rccl_merge|bitmorph|rccl_final|hyst_sweep' \
    --launch-skip 40 -c 14 -o gpurun_out/s3_top $B --steps 2 --warmup 3 > gpurun_out/s3_ncu2.log 2>&1
tail -1 gpurun_out/s3_ncu2.log
