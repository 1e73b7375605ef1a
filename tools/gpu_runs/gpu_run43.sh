#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r43_full_gpu_suite.log 2>&1 ) 2> gpurun_out/r43_time.txt; tail -3 gpurun_out/r43_full_gpu_suite.log | cut -c1-200; grep real gpurun_out/r43_time.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r43_smoke.log 2>&1; tail -1 gpurun_out/r43_smoke.log | cut -c1-200
( time timeout 900 python bench.py --impl reference > gpurun_out/r43_bench_reference.json 2> gpurun_out/r43_bench_reference.err ) 2> gpurun_out/r43_time_ref.txt; cut -c1-300 gpurun_out/r43_bench_reference.json; grep real gpurun_out/r43_time_ref.txt
( time timeout 900 python bench.py > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err ) 2> gpurun_out/r43_time_bench.txt; python -c "
import json; d=json.loads(open('gpurun_out/r43_bench.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline'])"; grep real gpurun_out/r43_time_bench.txt
