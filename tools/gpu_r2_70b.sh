#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29950 bench.py --gpus 8 --steps 2 --warmup 3 --model llama3_70b --no-e2e > gpurun_out/r2_cfg8_llama70b.json 2> gpurun_out/r2_cfg8_llama70b.err; echo "llama70b rc=$?"; tail -1 gpurun_out/r2_cfg8_llama70b.json | cut -c1-1500; grep -E "Error|error" gpurun_out/r2_cfg8_llama70b.err | head -5
