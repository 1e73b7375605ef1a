#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_targets.py tests/test_gpu_eval.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/r26_tests.log 2>&1; grep -E "assign_targets|passed|failed|Error|rror" gpurun_out/r26_tests.log | tail -12 | cut -c1-250
