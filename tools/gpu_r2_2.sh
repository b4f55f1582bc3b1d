#!/bin/bash
# 2 GPUs: bench both arms at N=2 (loss parity on fresh data per step)
mkdir -p gpurun_out
P=29600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n2.json | cut -c1-3000; tail -5 gpurun_out/r2_bench_n2.err | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref_n2.json | cut -c1-2000; tail -3 gpurun_out/r2_bench_ref_n2.err | cut -c1-600
