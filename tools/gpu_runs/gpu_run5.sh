#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py tests/test_gpu_boxes.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r5.log; tail -25 gpurun_out/pytest_r5.log | cut -c1-300
timeout 300 python tools/profile_layers.py > gpurun_out/layers.txt 2>&1; cp gpurun_out/layers.csv gpurun_out/layers_clean.csv; head -12 gpurun_out/layers.txt; tail -1 gpurun_out/layers.txt
for B in 1 4; do
  timeout 900 python bench.py --steps 24 --warmup 4 --scenes-per-step $B --skip-cpu-baseline > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "bench B=$B exit $?"; tail -2 gpurun_out/bench_b$B.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_b$B.json'));print('B=$B value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3),'clk',d['clocks'])"
done
timeout 900 python tools/nms_sweep.py > gpurun_out/nms_sweep.log 2>&1; tail -14 gpurun_out/nms_sweep.log
