#!/bin/bash
# MoE with expert parallelism over NVLink: tokens are pushed to / pulled from the owning rank's symmetric heap by our dispatch kernels ("flex" dispatcher).
set -euo pipefail
cd "$(dirname "$0")/.."
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 pretrain_gpt.py \
  --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --group-query-attention --num-query-groups 8 --ffn-hidden-size 14336 \
  --num-experts 8 --moe-router-topk 2 --moe-grouped-gemm --moe-token-dispatcher-type flex --expert-model-parallel-size 8 \
  --moe-aux-loss-coeff 0.01 --moe-router-load-balancing-type aux_loss \
  --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope --untie-embeddings-and-output-weights --bf16 \
  --seq-length 4096 --max-position-embeddings 4096 --micro-batch-size 1 --global-batch-size 64 --train-iters 100 --lr 1e-4 \
  --use-distributed-optimizer --mock-data --tokenizer-type NullTokenizer --vocab-size 32000 --log-interval 10 "$@"
