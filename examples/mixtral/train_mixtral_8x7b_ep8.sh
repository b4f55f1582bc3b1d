#!/bin/bash
# Mixtral 8x7B: 8 experts, top-2, expert parallel 8 through the NVLink push / pull dispatcher ("flex") and the grouped tcgen05 GEMM — BASELINE config #4.
source "$(dirname "$0")/../_common.sh"
PAR="--expert-model-parallel-size 8 --moe-token-dispatcher-type flex"
if [ "${TINY:-0}" = "1" ]; then PAR="--moe-token-dispatcher-type alltoall"; fi
$LAUNCH "$ROOT/pretrain_gpt.py" --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --num-query-groups 8 --ffn-hidden-size 14336 --seq-length 4096 --max-position-embeddings 4096 \
  --num-experts 8 --moe-router-topk 2 --moe-grouped-gemm --moe-aux-loss-coeff 1e-2 --moe-router-load-balancing-type aux_loss $PAR \
  --micro-batch-size 1 --global-batch-size 8 --train-iters 100 --lr 1e-4 --bf16 --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope \
  --untie-embeddings-and-output-weights --use-distributed-optimizer $DATA $TOK --vocab-size 32000 --log-interval 10 $TINY_ARGS "$@"
