#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r29_full_gpu_suite.log 2>&1 ) 2> gpurun_out/r29_time.txt; tail -4 gpurun_out/r29_full_gpu_suite.log | cut -c1-200; grep real gpurun_out/r29_time.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r29_smoke.log 2>&1; tail -3 gpurun_out/r29_smoke.log | cut -c1-200
