#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_boxes.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r17_tests.log 2>&1; tail -3 gpurun_out/r17_tests.log | cut -c1-300
timeout 600 python tools/profile_layers.py anchor 1 > gpurun_out/r17_layers_anchor_B1.txt 2>&1; grep "post" gpurun_out/r17_layers_anchor_B1.txt | head -3
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r17_bench.log 2>&1; tail -1 gpurun_out/r17_bench.log | cut -c1-400
