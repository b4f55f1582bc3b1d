#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x ) > gpurun_out/r2_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/r2_gemm_tests.log | cut -c1-300
timeout 200 python tools/gemm_bench.py qkv_tp8 proj_tp8 fc1_tp8 fc2_tp8 proj_tp1 fc2_tp1 > gpurun_out/r2_gemm_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r2_gemm_bench.log | cut -c1-900
