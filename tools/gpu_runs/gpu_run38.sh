#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_swin.py tests/test_gpu_fcos.py tests/test_gpu_e2e.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/r38_tests.log 2>&1; grep -E "\[fp16\]|passed|failed|rror" gpurun_out/r38_tests.log | tail -40 | cut -c1-200
