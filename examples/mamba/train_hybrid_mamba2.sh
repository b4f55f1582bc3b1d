#!/bin/bash
# Hybrid Mamba-2 / attention / MLP stack (pattern symbols: M mamba, * attention, - MLP).
source "$(dirname "$0")/../_common.sh"
SSM="--mamba-state-dim 128 --mamba-head-dim 64 --mamba-num-groups 8 --hybrid-override-pattern M-M-M*-M-M-M*-M-M-M*-M-M-M*-"
if [ "${TINY:-0}" = "1" ]; then SSM="--mamba-state-dim 16 --mamba-head-dim 16 --mamba-num-groups 1 --hybrid-override-pattern M*"; fi
$LAUNCH "$ROOT/pretrain_mamba.py" --num-layers 24 --hidden-size 2048 --num-attention-heads 16 --ffn-hidden-size 8192 --seq-length 4096 --max-position-embeddings 4096 \
  --micro-batch-size 1 --global-batch-size 8 --train-iters 100 --lr 3e-4 --bf16 --normalization RMSNorm --disable-bias-linear $DATA $TOK --vocab-size 50304 --log-interval 10 $TINY_ARGS $SSM "$@"
