#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r22_tests.log 2>&1; tail -12 gpurun_out/r22_tests.log | cut -c1-250
timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r22_layers_anchor_B4.txt 2>&1; grep -E "c3\+res|lat3|L0.ds|L0.c1|sum of" gpurun_out/r22_layers_anchor_B4.txt | head -14 | cut -c1-160
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r22_bench.log 2>&1; tail -1 gpurun_out/r22_bench.log | cut -c1-200
