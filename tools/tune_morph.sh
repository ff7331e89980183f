#!/bin/bash
# times the fused morphology kernel for several band heights (ms per 50 pages) and checks parity once
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_primitives.py -m gpu -q -x -k "config1 or config2 or odd_sizes or edge_cases or pipeline" 2>&1 | tail -2
for th in 48 64 80 96; do
SYNSEG_MORPH_TH=$th python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('TH $th: bit_dilate_erode', k['bit_dilate_erode']['ms_per_step'], 'step', round(d['ms_per_step'],4), 'serial', round(d['roofline']['serial_step_ms'],4))"
done
