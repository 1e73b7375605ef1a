#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/r28_eval.log 2>&1; grep -E "recall top|passed|failed|rror" gpurun_out/r28_eval.log | tail -10 | cut -c1-220
timeout 600 python bench.py --skip-cpu-baseline --precision fp16 > gpurun_out/r28_bench_fp16.log 2>&1; tail -1 gpurun_out/r28_bench_fp16.log | cut -c1-260
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r28_bench_bf16.log 2>&1; tail -1 gpurun_out/r28_bench_bf16.log | cut -c1-260
