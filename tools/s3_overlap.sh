#!/bin/bash
# page chunks x streams of synseg_detect_pages with the current kernels (ms per 50-page step, timed region)
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e"
for v in ${@:-"SYNSEG_OVERLAP=2,SYNSEG_STREAMS=2" "SYNSEG_OVERLAP=3,SYNSEG_STREAMS=3"}; do
  env ${v//,/ } $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('[$v] step', round(d['ms_per_step'],4), 'pages/s', round(d['value']))"
done
