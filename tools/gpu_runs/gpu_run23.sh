#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for t in stem l0c2; do
NCU_TARGET=$t timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/r23_prof_$t -f python tools/ncu_target.py > gpurun_out/r23_ncu_$t.log 2>&1
tail -1 gpurun_out/r23_ncu_$t.log | cut -c1-200
done
