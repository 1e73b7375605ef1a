#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -6 gpurun_out/other_configs.log | cut -c1-300
