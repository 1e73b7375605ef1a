#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_conv.py -m gpu -q -s -x --timeout 900 -p no:cacheprovider > gpurun_out/r27_tests.log 2>&1; grep -E "fp16 mode|passed|failed|Error|rror" gpurun_out/r27_tests.log | tail -12 | cut -c1-300
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r27_bench.log 2>&1; tail -1 gpurun_out/r27_bench.log | cut -c1-200
