#!/bin/bash
# T5-base style encoder-decoder span-corruption pre-training (encoder 512 tokens, decoder 128).
source "$(dirname "$0")/../_common.sh"
DEC="--decoder-seq-length 128"
if [ "${TINY:-0}" = "1" ]; then DEC="--decoder-seq-length 32"; fi
$LAUNCH "$ROOT/pretrain_t5.py" --num-layers 12 --hidden-size 768 --num-attention-heads 12 --ffn-hidden-size 3072 --seq-length 512 --max-position-embeddings 512 \
  --micro-batch-size 4 --global-batch-size 32 --train-iters 100 --lr 1e-4 --weight-decay 1e-2 --clip-grad 1.0 --bf16 $DATA $TOK --vocab-size 32128 --log-interval 10 $TINY_ARGS $DEC "$@"
