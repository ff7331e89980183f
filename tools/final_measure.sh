#!/bin/bash
# Round-end measurements on one B200: GPU tests, both bench arms, the ncu launch list and one full capture of the top kernels.
# usage: tools/final_measure.sh <tag>    (outputs under gpurun_out/<tag>_*)
cd "$(dirname "$0")/.."
t=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${t}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${t}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${t}_ref.json 2> gpurun_out/${t}_ref.err
python bench.py > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; cut -c1-300 gpurun_out/${t}_bench.json
# profiles: every kernel alone on one stream, 50 pages per launch (SYNSEG_OVERLAP=1), as in bench.py's profiled steps
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-dense --no-corpus --crops 0"
SYNSEG_OVERLAP=1 $B > gpurun_out/${t}_plain.log 2>&1 &&
SYNSEG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${t}_launches.csv $B > gpurun_out/${t}_ncu1.log 2>&1
SYNSEG_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:'canny_rgb|adaptive_mean|rccl_merge|bitmorph|rccl_final|hyst_sweep' \
    --launch-skip 40 -c 14 -f -o gpurun_out/${t}_top $B > gpurun_out/${t}_ncu2.log 2>&1
tail -1 gpurun_out/${t}_ncu2.log
