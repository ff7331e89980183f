#!/bin/bash
# A/B of library variants in one call (same box): tools/s3_ab.sh <tag> ...   ("" = the product library); two rounds to expose run-to-run noise
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu --no-corpus --crops 0 --no-e2e"
P=$PWD/synapta_image_segmentation_b200
for round in 1 2; do
for t in "$@"; do
  v=""; [ -n "$t" ] && [ "$t" != base ] && v="SYNSEG_LIB=$P/libsynseg_$t.so"
  case "$t" in env:*) v="${t#env:}";; esac          # env:NAME=VALUE runs the product library with that variable set
  env $v $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('[$t] step', round(d['ms_per_step'],4), 'dense', round(d['dense_pages']['ms_per_step'],4), {n:k[n]['ms_per_step'] for n in k if k[n]['ms_per_step'] > 0.03})"
done
done
