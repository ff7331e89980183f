#!/bin/bash
# session-3 experiment 3: local-prefix threshold test, cp.async prologue, 32-bit TMA offsets; variants: ring depth 3, register prologue, band heights
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py -m gpu -q -x -k "morph or adaptive or config1 or config2 or odd_sizes or edge_cases or ragged or variant or grey or gray" 2>&1 | tail -3
B="python bench.py --no-cpu --no-corpus --crops 0 --no-e2e"
P=$PWD/synapta_image_segmentation_b200
for v in "" "SYNSEG_LIB=$P/libsynseg_cr3.so" "SYNSEG_LIB=$P/libsynseg_pro8.so" "SYNSEG_TUNE_CANNY_BAND=32" "SYNSEG_TUNE_CANNY_BAND=64"; do
  env $v $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('[${v##*/}] step', round(d['ms_per_step'],4), 'dense', round(d['dense_pages']['ms_per_step'],4), {n:k[n]['ms_per_step'] for n in ('canny_rgb','adaptive_mean','bitmorph_h','bitmorph_v') if n in k})"
done
