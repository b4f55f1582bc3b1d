#!/bin/bash
# Pipeline 2 x tensor 4, interleaved 1F1B with 2 virtual stages per rank.
set -euo pipefail
cd "$(dirname "$0")/.."
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 pretrain_gpt.py \
  --model llama3_8b --tensor-model-parallel-size 4 --pipeline-model-parallel-size 2 --num-layers-per-virtual-pipeline-stage 8 --sequence-parallel --bf16 \
  --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 32 --train-iters 100 --lr 3e-4 \
  --use-distributed-optimizer --overlap-p2p-communication --mock-data --tokenizer-type NullTokenizer --vocab-size 128255 --log-interval 10 "$@"
