#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NRPN_PDL=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r31_bench_nopdl.log 2>&1; tail -1 gpurun_out/r31_bench_nopdl.log | cut -c1-200
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r31_bench_pdl.log 2>&1; tail -1 gpurun_out/r31_bench_pdl.log | cut -c1-200
NRPN_PDL=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r31_bench_nopdl2.log 2>&1; tail -1 gpurun_out/r31_bench_nopdl2.log | cut -c1-200
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r31_bench_pdl2.log 2>&1; tail -1 gpurun_out/r31_bench_pdl2.log | cut -c1-200
timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r31_layers_anchor_B4.txt 2>&1; grep -E "maxpool|pack_stem|sum of" gpurun_out/r31_layers_anchor_B4.txt | cut -c1-120
