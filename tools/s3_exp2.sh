#!/bin/bash
# session-3 experiment 2: 128-byte TMA units, clean morphology kernels; parity subset + A/B bench lines
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py -m gpu -q -x --durations=6 -k "morph or adaptive or config1 or config2 or odd_sizes or edge_cases or ragged or golden or variant or grey or gray" 2>&1 | tail -12
B="python bench.py --no-cpu --no-corpus --crops 0 --no-e2e"
for v in "" "SYNSEG_TUNE_AD_BAND=96" "SYNSEG_TUNE_AD_BAND=72" "SYNSEG_NO_TMA=1"; do
  env $v $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('[$v] step', round(d['ms_per_step'],4), 'dense', round(d['dense_pages']['ms_per_step'],4), {n:k[n]['ms_per_step'] for n in ('canny_rgb','adaptive_mean','bitmorph_h','bitmorph_v') if n in k})"
done
