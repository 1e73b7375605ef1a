#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_swin.py tests/test_gpu_fcos.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r40_tests.log 2>&1; tail -3 gpurun_out/r40_tests.log | cut -c1-200
for i in 1 2; do timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r40_bench_$i.log 2>&1; tail -1 gpurun_out/r40_bench_$i.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
