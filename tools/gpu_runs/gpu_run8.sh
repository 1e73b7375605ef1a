#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_swin.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r8.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r8.log; grep -E "swin config|passed|failed|exit|Error|assert|err" gpurun_out/pytest_r8.log | tail -30 | cut -c1-260
