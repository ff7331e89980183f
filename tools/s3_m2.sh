#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SYNSEG_LIB=$PWD/synapta_image_segmentation_b200/libsynseg_m2.so python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pipeline.py tests/test_gpu_properties.py -m gpu -q -x -k "ccl or canny or hyst or connected or config1 or config2 or edge_cases or pipeline_property" 2>&1 | tail -2
tools/s3_ab.sh base m2 mm6 mg16
