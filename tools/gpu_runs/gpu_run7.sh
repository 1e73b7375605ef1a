#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fcos.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r7.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r7.log; grep -E "fcos_small|passed|failed|exit|Error|assert" gpurun_out/pytest_r7.log | tail -25 | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_e2e.py -m gpu -q --timeout 900 -p no:cacheprovider -k "rpn or small or padded" 2>&1 | tail -3
