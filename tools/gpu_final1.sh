#!/bin/bash
# 1-GPU final validation: full GPU test suite, smoke(), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/final_gpu_tests.log | cut -c1-300
timeout 600 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/final_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/final_bench_n1.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/final_bench_n1.log | cut -c1-1800 || tail -30 gpurun_out/final_bench_n1.log
