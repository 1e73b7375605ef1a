#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider -s > gpurun_out/r14_tests.log 2>&1; grep -E "stem 160|passed|failed|error" gpurun_out/r14_tests.log | tail -5 | cut -c1-300
timeout 600 python tools/profile_layers.py anchor 1 > gpurun_out/r14_layers_anchor_B1.txt 2>&1; head -14 gpurun_out/r14_layers_anchor_B1.txt | cut -c1-200; tail -1 gpurun_out/r14_layers_anchor_B1.txt
timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r14_layers_anchor_B4.txt 2>&1; head -30 gpurun_out/r14_layers_anchor_B4.txt | cut -c1-200; tail -1 gpurun_out/r14_layers_anchor_B4.txt
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r14_bench.log 2>&1; tail -1 gpurun_out/r14_bench.log | cut -c1-1200
