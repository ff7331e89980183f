#!/bin/bash
# morphology passes: 4 vs 8 output words per thread in the row pass, threads per CTA of the van Herk column pass (ms per 50 pages)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_properties.py tests/test_gpu_pipeline.py -m gpu -q -x -k "morph or pipeline or config or ragged or hints_match_reference_golden" 2>&1 | tail -2
for v in "0 256" "1 256" "0 128" "0 64"; do
  set -- $v
  env_h=""; [ $1 = 1 ] && env_h="SYNSEG_MORPH_H4=1"
  env $env_h SYNSEG_MORPH_VT=$2 python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('row pass 4-word=$1 column T=$2: bitmorph_h', k['bitmorph_h']['ms_per_step'], 'bitmorph_v', k['bitmorph_v']['ms_per_step'], 'step', round(d['ms_per_step'],4))"
done
