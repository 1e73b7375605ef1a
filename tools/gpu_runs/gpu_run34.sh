#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_conv.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r34_tests.log 2>&1; tail -6 gpurun_out/r34_tests.log | cut -c1-250
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r34_bench_dataset.log 2>&1; tail -1 gpurun_out/r34_bench_dataset.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dataset layout', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
timeout 600 python bench.py --skip-cpu-baseline --input-layout ncdhw > gpurun_out/r34_bench_ncdhw.log 2>&1; tail -1 gpurun_out/r34_bench_ncdhw.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ncdhw layout', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
