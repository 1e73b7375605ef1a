#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py tests/test_gpu_swin.py tests/test_gpu_fcos.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r30_tests.log 2>&1; tail -2 gpurun_out/r30_tests.log | cut -c1-200
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r30_bench_pdl.log 2>&1; tail -1 gpurun_out/r30_bench_pdl.log | cut -c1-200
NRPN_PDL=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r30_bench_nopdl.log 2>&1; tail -1 gpurun_out/r30_bench_nopdl.log | cut -c1-200
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -3 gpurun_out/other_configs.log | cut -c1-200
NRPN_PDL=0 timeout 900 python tools/bench_configs.py > gpurun_out/other_configs_nopdl.log 2>&1; tail -3 gpurun_out/other_configs_nopdl.log | cut -c1-200
