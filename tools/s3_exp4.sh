#!/bin/bash
# session-3 experiment 4: background stats by complement, empty-segment fast paths, batched border loads in the column morphology
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py tests/test_gpu_properties.py tests/test_gpu_regions.py -m gpu -q -x -k "not golden and not 591 and not exhaustive" 2>&1 | tail -3
B="python bench.py --no-cpu --no-corpus --crops 0 --no-e2e"
for v in "" ; do
  env $v $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('[${v##*/}] step', round(d['ms_per_step'],4), 'dense', round(d['dense_pages']['ms_per_step'],4), {n:k[n]['ms_per_step'] for n in k})"
done
