#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 ./tools/micro/mma_floor 2>&1 | tee gpurun_out/r12_mma_floor.txt
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "slab or stem" > gpurun_out/r12_conv_tests.log 2>&1; tail -8 gpurun_out/r12_conv_tests.log | cut -c1-300
