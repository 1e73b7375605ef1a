#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py tests/test_gpu_fcos.py tests/test_gpu_swin.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r20_tests.log 2>&1; tail -3 gpurun_out/r20_tests.log | cut -c1-300
timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r20_layers_anchor_B4.txt 2>&1; tail -1 gpurun_out/r20_layers_anchor_B4.txt
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r20_bench.log 2>&1; tail -1 gpurun_out/r20_bench.log | cut -c1-300
