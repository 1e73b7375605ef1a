#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/r11_conv_tests.log 2>&1; tail -25 gpurun_out/r11_conv_tests.log | cut -c1-300
timeout 600 python tools/profile_layers.py anchor 1 > gpurun_out/r11_layers_anchor_B1.txt 2>&1; head -12 gpurun_out/r11_layers_anchor_B1.txt | cut -c1-200; tail -1 gpurun_out/r11_layers_anchor_B1.txt
timeout 600 python tools/profile_layers.py resnet50_fcos_160x256x256 1 > gpurun_out/r11_layers_fcos_B1.txt 2>&1; head -16 gpurun_out/r11_layers_fcos_B1.txt | cut -c1-200; tail -1 gpurun_out/r11_layers_fcos_B1.txt
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r11_bench.log 2>&1; tail -1 gpurun_out/r11_bench.log | cut -c1-900
