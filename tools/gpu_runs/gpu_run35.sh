#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
# (1) launch list of the bench command itself
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r35_bench_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r35_bench_under_ncu.log 2>&1
python tools/ncu_summarize.py gpurun_out/r35_bench_launches.csv "bench.py --steps 2 --warmup 1 (4 scenes per step), launch list" > gpurun_out/r35_bench_launches.md 2> gpurun_out/r35_sum.err; head -14 gpurun_out/r35_bench_launches.md | cut -c1-150; tail -1 gpurun_out/r35_bench_launches.md
# (2) full captures of the dominant kernel (head layer), the slab stem, a TMA-epilogue 1^3 layer, window attention is in config 3
for t in head stem l0c3; do
NCU_TARGET=$t timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/r35_prof_$t -f python tools/ncu_target.py > gpurun_out/r35_ncu_$t.log 2>&1
done
python tools/ncu_full_summary.py gpurun_out/r35_ncu_full_summary.json head=gpurun_out/r35_prof_head.ncu-rep stem=gpurun_out/r35_prof_stem.ncu-rep l0c3=gpurun_out/r35_prof_l0c3.ncu-rep > gpurun_out/r35_summary.log 2>&1; tail -5 gpurun_out/r35_summary.log | cut -c1-200
