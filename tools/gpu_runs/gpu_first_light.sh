#!/bin/bash
# GPU bring-up: box kernels first (all tests), then the conv kernel (stop at first failure), full logs kept.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_boxes.py -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/pytest_boxes.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_boxes.log
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_conv.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -40 gpurun_out/pytest_boxes.log; tail -80 gpurun_out/pytest_conv.log
