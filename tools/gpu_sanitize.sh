#!/bin/bash
# compute-sanitizer passes: memcheck on every kernel family, racecheck on the shared-memory kernels.  Tight timeouts: the GPU budget is nearly spent.
mkdir -p gpurun_out
export MEGATRON_B200_GEMM=tcgen05
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_ops.py --group simple > gpurun_out/sanitize_memcheck_simple.log 2>&1; echo "memcheck simple rc=$?"; grep -E "ERROR SUMMARY|Invalid|group ok" gpurun_out/sanitize_memcheck_simple.log | tail -4
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_ops.py --group tensor_core > gpurun_out/sanitize_memcheck_tc.log 2>&1; echo "memcheck tensor-core rc=$?"; grep -E "ERROR SUMMARY|Invalid|group ok" gpurun_out/sanitize_memcheck_tc.log | tail -4
timeout 60 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python tools/sanitize_ops.py --group simple > gpurun_out/sanitize_racecheck_simple.log 2>&1; echo "racecheck simple rc=$?"; grep -E "RACECHECK SUMMARY|hazard|group ok" gpurun_out/sanitize_racecheck_simple.log | tail -4
