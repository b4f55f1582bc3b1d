#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/gemm2.log 2>&1; echo "gemm tests rc=$?"; tail -4 gpurun_out/gemm2.log
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench2.log 2>&1; grep '^{' gpurun_out/gemm_bench2.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_full2.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_full2.log || tail -20 gpurun_out/bench_full2.log
