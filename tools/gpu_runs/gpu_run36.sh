#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r36_memcheck.log python -m pytest tests/test_gpu_conv.py tests/test_gpu_targets.py tests/test_gpu_eval.py -m gpu -q -x --timeout 1400 -p no:cacheprovider -k "not full_size and not faster and not large_p2 and not within_half" > gpurun_out/r36_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r36_memcheck_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r36_memcheck.log | head -10
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r36_memcheck_swin.log python -m pytest tests/test_gpu_swin.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "window_attention or layernorm or patch_embed" > gpurun_out/r36_memcheck_swin_pytest.log 2>&1; echo "memcheck swin rc=$?"; tail -2 gpurun_out/r36_memcheck_swin_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r36_memcheck_swin.log | head -10
