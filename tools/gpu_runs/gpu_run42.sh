#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/profile_layers.py config3_swin_s_fcos_200x200x130 1 > gpurun_out/r42_layers_config3.txt 2>&1; tail -1 gpurun_out/r42_layers_config3.txt
timeout 600 python tools/profile_layers.py config1_vgg19_anchor_32 1 > gpurun_out/r42_layers_config1.txt 2>&1; head -12 gpurun_out/r42_layers_config1.txt | cut -c1-150; tail -1 gpurun_out/r42_layers_config1.txt
