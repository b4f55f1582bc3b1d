#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log | grep '^{'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 6 -c 2 -o gpurun_out/prof_gemm python tools/gemm_bench.py fc1_tp1 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
