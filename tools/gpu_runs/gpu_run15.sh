#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 900 -p no:cacheprovider -s -k "vgg" > gpurun_out/r15_vgg.log 2>&1; grep -E "vgg config|passed|failed" gpurun_out/r15_vgg.log | tail -12 | cut -c1-200
timeout 600 python tools/profile_layers.py config3_swin_s_fcos_200x200x130 1 > gpurun_out/r15_layers_config3_B1.txt 2>&1; head -45 gpurun_out/r15_layers_config3_B1.txt | cut -c1-200; tail -1 gpurun_out/r15_layers_config3_B1.txt
