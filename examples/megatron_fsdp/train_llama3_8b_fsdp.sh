#!/bin/bash
# ZeRO-3 style fully-sharded data parallel (parameters, gradients and optimizer state sharded over the data-parallel group; TP = PP = 1).
source "$(dirname "$0")/../_common.sh"
$LAUNCH "$ROOT/pretrain_gpt.py" --model llama3_8b --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 8 --train-iters 100 --lr 3e-4 --bf16 \
  --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope --untie-embeddings-and-output-weights \
  --use-megatron-fsdp --data-parallel-sharding-strategy optim_grads_params $DATA $TOK --vocab-size 128256 --log-interval 10 $TINY_ARGS "$@"
