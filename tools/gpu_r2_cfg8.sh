#!/bin/bash
# 8 GPUs: the other three BASELINE.json configs (repo arm) + the one-MoE-layer benchmark
mkdir -p gpurun_out
P=29900
run() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 3 --warmup 3 "$@" > gpurun_out/r2_cfg8_$name.json 2> gpurun_out/r2_cfg8_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_cfg8_$name.json | cut -c1-1500; grep -E "Error|error|Traceback" gpurun_out/r2_cfg8_$name.err | head -5; tail -3 gpurun_out/r2_cfg8_$name.err | cut -c1-300; P=$((P+1)); }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29890 tools/moe_layer_bench.py > gpurun_out/r2_moe_layer.log 2>&1; echo "moe layer rc=$?"; grep moe_layer gpurun_out/r2_moe_layer.log; tail -3 gpurun_out/r2_moe_layer.log | cut -c1-300
run gpt67 --model gpt3_6.7b
run mixtral_flex --model mixtral_8x7b --dispatcher flex
run llama70b --model llama3_70b --no-e2e
