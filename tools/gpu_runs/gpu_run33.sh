#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/r33_eval.log 2>&1; grep -E "AP@|passed|failed|rror" gpurun_out/r33_eval.log | tail -10 | cut -c1-220
