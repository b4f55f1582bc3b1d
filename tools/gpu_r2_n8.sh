#!/bin/bash
# 8 GPUs: our arm (no-SP primary + SP variant + pair-op self check), then the multi-GPU tests
mkdir -p gpurun_out
P=29700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n8.json | cut -c1-3500; tail -5 gpurun_out/r2_bench_n8.err | cut -c1-600
( time timeout 300 python -m pytest tests/test_nvlink_gpu.py -q -x -s ) > gpurun_out/r2_mg_tests_n8.log 2>&1; echo "mg tests rc=$?"; tail -4 gpurun_out/r2_mg_tests_n8.log | cut -c1-1500
