#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_boxes.py tests/test_gpu_fcos.py -m gpu -q -x --timeout 600 -p no:cacheprovider -s > gpurun_out/r25_tests.log 2>&1; grep -E "swin config|passed|failed|Error|err " gpurun_out/r25_tests.log | tail -12 | cut -c1-250
NRPN_ATTN_TC=0 timeout 600 python tools/profile_layers.py config3_swin_s_fcos_200x200x130 1 > gpurun_out/r25_layers_config3_cudacore.txt 2>&1; grep -E "window_attention" gpurun_out/r25_layers_config3_cudacore.txt | head -4 | cut -c1-160
timeout 600 python tools/profile_layers.py config3_swin_s_fcos_200x200x130 1 > gpurun_out/r25_layers_config3_tc.txt 2>&1; grep -E "window_attention| post " gpurun_out/r25_layers_config3_tc.txt | head -5 | cut -c1-160
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -3 gpurun_out/other_configs.log | cut -c1-260
