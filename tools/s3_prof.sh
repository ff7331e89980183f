#!/bin/bash
# full ncu capture (with source) of the kernels named in $1 (regex), one profiled 50-page step
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu --no-dense --no-corpus --crops 0 --no-e2e --steps 2 --warmup 3"
SYNSEG_OVERLAP=1 $B > gpurun_out/${2:-s3}_plain.log 2>&1 &&
SYNSEG_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:"$1" --launch-skip ${3:-30} -c ${4:-6} -o gpurun_out/${2:-s3}_top -f $B > gpurun_out/${2:-s3}_ncu.log 2>&1
tail -1 gpurun_out/${2:-s3}_ncu.log
