#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/attn_probe.py > gpurun_out/attn_probe.log 2>&1
cat gpurun_out/attn_probe.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"
grep '^{' gpurun_out/bench_full.log || tail -20 gpurun_out/bench_full.log
MEGATRON_B200_GEMM=cublas timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_full_cublas.log 2>&1; echo "bench full cublas rc=$?"
grep '^{' gpurun_out/bench_full_cublas.log || tail -20 gpurun_out/bench_full_cublas.log
timeout 1500 python bench.py --impl reference --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
grep '^{' gpurun_out/bench_ref.log || tail -30 gpurun_out/bench_ref.log
