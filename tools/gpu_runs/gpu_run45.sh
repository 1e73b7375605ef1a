#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference > gpurun_out/r45_bench_reference.json 2> gpurun_out/r45_bench_reference.err ) 2> gpurun_out/r45_time_ref.txt; cut -c1-120 gpurun_out/r45_bench_reference.json; python -c "
import json; d=json.loads(open('gpurun_out/r45_bench_reference.json').read().strip().splitlines()[-1]); print(d['cpu_baseline'])"; grep real gpurun_out/r45_time_ref.txt
( time timeout 900 python bench.py > gpurun_out/r45_bench.json 2> gpurun_out/r45_bench.err ) 2> gpurun_out/r45_time_bench.txt; python -c "
import json; d=json.loads(open('gpurun_out/r45_bench.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}); print(d['e2e']['value']); print(d['cpu_baseline'])"; grep real gpurun_out/r45_time_bench.txt
