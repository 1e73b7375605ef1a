#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NRPN_SIDE_PRIORITY=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r32_bench_prio0.log 2>&1; tail -1 gpurun_out/r32_bench_prio0.log | cut -c1-120
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r32_bench_prio1.log 2>&1; tail -1 gpurun_out/r32_bench_prio1.log | cut -c1-120
NRPN_SIDE_PRIORITY=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r32_bench_prio0b.log 2>&1; tail -1 gpurun_out/r32_bench_prio0b.log | cut -c1-120
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r32_bench_prio1b.log 2>&1; tail -1 gpurun_out/r32_bench_prio1b.log | cut -c1-120
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -3 gpurun_out/other_configs.log | cut -c1-160
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r32_tests.log 2>&1; tail -2 gpurun_out/r32_tests.log | cut -c1-200
