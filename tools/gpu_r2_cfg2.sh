#!/bin/bash
# 2 GPUs: debug the PP / interleaved-PP / EP / DP paths on tiny models before the 8-GPU runs
mkdir -p gpurun_out
P=29800
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e "$@" > gpurun_out/r2_cfg_$name.json 2> gpurun_out/r2_cfg_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_cfg_$name.json | cut -c1-700; grep -E "Error|error|Traceback" gpurun_out/r2_cfg_$name.err | head -5; tail -4 gpurun_out/r2_cfg_$name.err | cut -c1-300; P=$((P+1)); }
run pp2 --model tiny_llama --tp 1 --pp 2 --global-batch 8
run pp2vp2 --model tiny_llama --tp 1 --pp 2 --vp 2 --layers 4 --global-batch 8
run ep2 --model tiny_mixtral --tp 1 --pp 1 --ep 2 --global-batch 8
run ep2flex --model tiny_mixtral --tp 1 --pp 1 --ep 2 --global-batch 8 --dispatcher flex
run dp2 --model tiny_llama --tp 1 --pp 1 --global-batch 8
run gpt67_l4 --model gpt3_6.7b --tp 1 --pp 2 --vp 2 --layers 4 --global-batch 4
run mixtral_l2 --model mixtral_8x7b --tp 1 --ep 2 --layers 2 --global-batch 2
