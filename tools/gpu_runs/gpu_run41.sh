#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r41_tests.log 2>&1; tail -2 gpurun_out/r41_tests.log | cut -c1-200
for m in 1 0; do NRPN_CONV_NARROW=$m timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r41_layers_B4_narrow$m.txt 2>&1; echo "narrow=$m"; grep -E "L3.c2|L2.c2|L3.c1|lat0|sum of" gpurun_out/r41_layers_B4_narrow$m.txt | head -8 | cut -c1-140; done
for m in 1 0; do NRPN_CONV_NARROW=$m timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r41_bench_narrow$m.log 2>&1; tail -1 gpurun_out/r41_bench_narrow$m.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench narrow=$m', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"; done
NRPN_CONV_NARROW=1 timeout 900 python tools/bench_configs.py > gpurun_out/other_configs.log 2>&1; tail -3 gpurun_out/other_configs.log | cut -c1-150
