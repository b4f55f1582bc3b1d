#!/bin/bash
# BERT-large style masked-LM + sentence-order pre-training.
source "$(dirname "$0")/../_common.sh"
$LAUNCH "$ROOT/pretrain_bert.py" --num-layers 24 --hidden-size 1024 --num-attention-heads 16 --seq-length 512 --max-position-embeddings 512 --micro-batch-size 4 --global-batch-size 32 \
  --train-iters 100 --lr 1e-4 --weight-decay 1e-2 --clip-grad 1.0 --bf16 $DATA $TOK --vocab-size 30592 --log-interval 10 $TINY_ARGS "$@"
