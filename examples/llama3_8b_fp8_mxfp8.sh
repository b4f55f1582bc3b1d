#!/bin/bash
# MXFP8 training recipe: every linear GEMM (fprop, dgrad, wgrad) runs block-scaled on the tensor cores; first and last layer stay bf16.
set -euo pipefail
cd "$(dirname "$0")/.."
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 pretrain_gpt.py \
  --model llama3_8b --tensor-model-parallel-size 2 --sequence-parallel --bf16 --fp8-format hybrid --fp8-recipe mxfp8 \
  --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 32 \
  --train-iters 100 --lr 3e-4 --use-distributed-optimizer --mock-data --tokenizer-type NullTokenizer --vocab-size 128255 --log-interval 10 "$@"
