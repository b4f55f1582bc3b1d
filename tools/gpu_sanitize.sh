#!/bin/bash
mkdir -p gpurun_out
export MEGATRON_B200_GEMM=tcgen05
timeout 50 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python tools/sanitize_ops.py --group all > gpurun_out/sanitize_synccheck_all.log 2>&1; echo "synccheck all rc=$?"; grep -E "ERROR SUMMARY|Barrier|group ok" gpurun_out/sanitize_synccheck_all.log | tail -4
timeout 50 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_ops.py --group simple > gpurun_out/sanitize_initcheck_simple.log 2>&1; echo "initcheck simple rc=$?"; grep -E "ERROR SUMMARY|Uninitialized|group ok" gpurun_out/sanitize_initcheck_simple.log | tail -4
