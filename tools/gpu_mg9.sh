#!/bin/bash
# 2-GPU sanity of the final tree: default bench (auto → fused TP kernels) exactly as the driver launches it
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_final_n2.log 2>&1; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_final_n2.log | cut -c1-1100 || tail -25 gpurun_out/bench_final_n2.log
