#!/bin/bash
# 2 GPUs: data-parallel gradient reduce-scatter / param all-gather over NVLink (default) vs NCCL, tiny model + 8-layer Llama-3-8B slice
mkdir -p gpurun_out
P=29980
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e "$@" > gpurun_out/r2_dp_$name.json 2> gpurun_out/r2_dp_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_dp_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['loss_by_step'], d['config']['parallelism'])"; grep -E "Error|error|Warning: NVLink" gpurun_out/r2_dp_$name.err | head -3; P=$((P+1)); }
run tiny_nvl --model tiny_llama --tp 1 --pp 1 --global-batch 8
MEGATRON_B200_DP_COMM=nccl run tiny_nccl --model tiny_llama --tp 1 --pp 1 --global-batch 8
run l8_nvl --model llama3_8b_dp --layers 8 --tp 1 --pp 1 --global-batch 4
MEGATRON_B200_DP_COMM=nccl run l8_nccl --model llama3_8b_dp --layers 8 --tp 1 --pp 1 --global-batch 4
