#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_fcos.py tests/test_gpu_e2e.py tests/test_gpu_swin.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r16_tests.log 2>&1; tail -5 gpurun_out/r16_tests.log | cut -c1-300
timeout 900 python tools/nms_sweep.py > gpurun_out/r16_nms_sweep.log 2>&1; tail -14 gpurun_out/r16_nms_sweep.log | cut -c1-200
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -3 gpurun_out/other_configs.log | cut -c1-300
timeout 600 python tools/profile_layers.py anchor 1 > gpurun_out/r16_layers_anchor_B1.txt 2>&1; grep "post" gpurun_out/r16_layers_anchor_B1.txt | head -3
