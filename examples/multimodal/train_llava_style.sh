#!/bin/bash
# LLaVA-style vision-language pre-training: CLIP-ViT tower -> projector -> language model (synthetic image / text pairs unless a dataset is given).
source "$(dirname "$0")/../_common.sh"
VIT="--img-h 336 --img-w 336 --patch-dim 14 --vision-num-layers 24 --vision-hidden-size 1024 --vision-num-attention-heads 16"
if [ "${TINY:-0}" = "1" ]; then VIT="--img-h 28 --img-w 28 --patch-dim 14 --vision-num-layers 1 --vision-hidden-size 32 --vision-num-attention-heads 2"; fi
$LAUNCH "$ROOT/pretrain_vlm.py" --num-layers 16 --hidden-size 2048 --num-attention-heads 16 --ffn-hidden-size 5632 --seq-length 1024 --max-position-embeddings 1024 \
  --micro-batch-size 1 --global-batch-size 8 --train-iters 50 --lr 1e-4 --bf16 --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope \
  --untie-embeddings-and-output-weights $DATA $TOK --vocab-size 32000 --log-interval 5 $TINY_ARGS $VIT "$@"
