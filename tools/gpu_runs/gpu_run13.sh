#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for t in stem l0c2; do
NCU_TARGET=$t timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/r13_prof_$t -f python tools/ncu_target.py > gpurun_out/r13_ncu_$t.log 2>&1
tail -3 gpurun_out/r13_ncu_$t.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
