#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --dtype mxfp8 --main-grads bf16 > gpurun_out/r2_bench_n1_mxfp8.json 2> gpurun_out/r2_bench_n1_mxfp8.err; echo "mxfp8 rc=$?"; tail -1 gpurun_out/r2_bench_n1_mxfp8.json | cut -c1-1000; grep -E "Error|error" gpurun_out/r2_bench_n1_mxfp8.err | head -5; tail -3 gpurun_out/r2_bench_n1_mxfp8.err | cut -c1-300
