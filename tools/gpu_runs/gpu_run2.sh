#!/bin/bash
# second GPU trip: end-to-end parity, smoke, first bench line, ncu launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_e2e.log
tail -30 gpurun_out/pytest_e2e.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
