#!/bin/bash
# third GPU trip: re-validate conv + e2e after the epilogue / multi-CTA / side-stream changes, per-layer table, bench,
# ncu launch list of graph replays, one full ncu capture of the dominant kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_r3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r3.log; tail -15 gpurun_out/pytest_r3.log
timeout 300 python tools/profile_layers.py > gpurun_out/layers.txt 2>&1; head -32 gpurun_out/layers.txt
timeout 900 python bench.py --steps 30 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 300 -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 4 > gpurun_out/bench_ncu.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm_kernel.*256 -s 8 -c 2 -o gpurun_out/prof_conv256 python tools/profile_layers.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
