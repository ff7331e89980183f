#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py tests/test_gpu_properties.py -m gpu -q -x -k "canny or hyst or config1 or config2 or odd_sizes or edge_cases or pipeline or variant" 2>&1 | tail -3
tools/s3_ab.sh base prev
