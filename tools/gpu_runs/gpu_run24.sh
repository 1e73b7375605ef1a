#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
for c in anchor config3_swin_s_fcos_200x200x130; do
timeout 900 ncu --metrics $M --clock-control none --nvtx --nvtx-include "step/" --csv --log-file gpurun_out/r24_ncu_step_$c.csv python tools/ncu_step.py $c > gpurun_out/r24_ncu_step_$c.log 2>&1
python tools/ncu_summarize.py gpurun_out/r24_ncu_step_$c.csv "$c" > gpurun_out/r24_ncu_step_$c.md 2>gpurun_out/r24_sum_$c.err; head -16 gpurun_out/r24_ncu_step_$c.md | cut -c1-200; tail -1 gpurun_out/r24_ncu_step_$c.md
done
timeout 600 python tools/profile_layers.py anchor 4 > gpurun_out/r24_layers_anchor_B4.txt 2>&1; grep -E "lat3|sum of" gpurun_out/r24_layers_anchor_B4.txt | head -3 | cut -c1-160
timeout 600 python bench.py > gpurun_out/r24_bench.log 2>&1; tail -1 gpurun_out/r24_bench.log > gpurun_out/r24_bench.json; cut -c1-300 gpurun_out/r24_bench.json
