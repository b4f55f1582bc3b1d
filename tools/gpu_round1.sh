#!/bin/bash
# First GPU pass: op numerics, GEMM numerics (own timeout: a bad descriptor hangs), smoke, short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/ops.log 2>&1; echo "ops rc=$?" >> gpurun_out/summary.txt
timeout 240 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/ops.log; tail -30 gpurun_out/gemm.log
MEGATRON_B200_GEMM=cublas timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_cublas.log 2>&1; echo "smoke(cublas) rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/smoke_cublas.log gpurun_out/smoke.log
MEGATRON_B200_GEMM=cublas timeout 600 python bench.py --layers 4 --steps 3 --warmup 2 > gpurun_out/bench_l4_cublas.log 2>&1; echo "bench l4 cublas rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --layers 4 --steps 3 --warmup 2 > gpurun_out/bench_l4.log 2>&1; echo "bench l4 rc=$?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_l4_cublas.log gpurun_out/bench_l4.log
cat gpurun_out/summary.txt
