#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py -m gpu -q -x -k "canny or config1 or config2 or odd_sizes or edge_cases or variant or golden" 2>&1 | tail -3
SYNSEG_LIB=$PWD/synapta_image_segmentation_b200/libsynseg_vg96.so python -m pytest tests/test_gpu_primitives.py tests/test_gpu_properties.py -m gpu -q -x -k "canny" 2>&1 | tail -2
tools/s3_ab.sh base vg0 vg96
