#!/bin/bash
# usage: tools/s3_run.sh "<pytest -k expression>" <variant tags ...>    quick parity subset, then the A/B of tools/s3_ab.sh
cd "$(dirname "$0")/.."
k="$1"; shift
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py -m gpu -q -x -k "$k" 2>&1 | tail -3
tools/s3_ab.sh "$@"
