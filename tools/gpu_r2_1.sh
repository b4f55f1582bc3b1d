#!/bin/bash
# round 2, call 1 (1 GPU): full GPU suite incl. the one-GPU multi-process IPC protocol tests, bench N=1 with fp32 main grads, attention baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_nvlink_ipc_gpu.py ) > gpurun_out/r2_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2_gpu_suite.log
( time timeout 600 python -m pytest tests/test_nvlink_ipc_gpu.py -q -s ) > gpurun_out/r2_ipc.log 2>&1; echo "ipc rc=$?"; tail -25 gpurun_out/r2_ipc.log | cut -c1-600
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n1.json | cut -c1-1500; tail -5 gpurun_out/r2_bench_n1.err | cut -c1-400
timeout 200 python tools/attn_probe.py > gpurun_out/r2_attn_probe.log 2>&1; echo "attn rc=$?"; tail -12 gpurun_out/r2_attn_probe.log | cut -c1-300
