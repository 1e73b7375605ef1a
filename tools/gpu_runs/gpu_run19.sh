#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for t in l0c3 lat3; do
NCU_TARGET=$t timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "target/" -o gpurun_out/r19_prof_$t -f python tools/ncu_target.py > gpurun_out/r19_ncu_$t.log 2>&1
tail -2 gpurun_out/r19_ncu_$t.log | cut -c1-200
done
