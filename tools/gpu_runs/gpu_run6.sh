#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r6.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r6.log; grep -E "vgg config|passed|failed|exit" gpurun_out/pytest_r6.log | tail -12
NRPN_SPLITK=1 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k split_k --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/profile_layers.py > gpurun_out/layers.txt 2>&1; cp gpurun_out/layers.csv gpurun_out/layers_clean.csv; tail -1 gpurun_out/layers.txt
for B in 1 2 4 8; do
  timeout 900 python bench.py --steps 16 --warmup 4 --scenes-per-step $B --skip-cpu-baseline > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "bench B=$B exit $?"; tail -2 gpurun_out/bench_b$B.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_b$B.json'));print('B=$B value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3),'clk',d['clocks'])"
done
