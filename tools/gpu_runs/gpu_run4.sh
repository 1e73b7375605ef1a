#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_boxes.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_r4.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r4.log; tail -25 gpurun_out/pytest_r4.log
timeout 300 python tools/profile_layers.py > gpurun_out/layers.txt 2>&1; cp gpurun_out/layers.csv gpurun_out/layers_clean.csv; head -14 gpurun_out/layers.txt; tail -1 gpurun_out/layers.txt
for B in 1 2 4; do
  timeout 900 python bench.py --steps 24 --warmup 4 --scenes-per-step $B > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "bench B=$B exit $?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_b$B.json'));print('B=$B value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3),'clk',d['clocks'])"
done
NCU_TARGET=head timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/ncu_head.log 2>&1; echo "ncu head exit $?"
NCU_TARGET=misc timeout 900 ncu --set full --clock-control none --nvtx --nvtx-include "target/" -o gpurun_out/prof_misc python tools/ncu_target.py > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc exit $?"
ls -la gpurun_out/*.ncu-rep
