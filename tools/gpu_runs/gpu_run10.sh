#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_swin.py tests/test_gpu_fcos.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r10_tests.log 2>&1; tail -15 gpurun_out/r10_tests.log | cut -c1-250
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -4 gpurun_out/other_configs.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r10_ref_arm.log 2>&1; tail -2 gpurun_out/r10_ref_arm.log | cut -c1-600
