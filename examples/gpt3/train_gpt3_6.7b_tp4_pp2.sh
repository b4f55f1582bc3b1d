#!/bin/bash
# GPT-3 6.7B (32 layers, h 4096, 32 heads, learned positions, GeLU, LayerNorm, biases, dropout 0.1), TP 4 x PP 2 with interleaved 1F1B — BASELINE config #3.
source "$(dirname "$0")/../_common.sh"
PAR="--tensor-model-parallel-size 4 --pipeline-model-parallel-size 2 --num-layers-per-virtual-pipeline-stage 8 --sequence-parallel"
if [ "${TINY:-0}" = "1" ]; then PAR=""; fi
$LAUNCH "$ROOT/pretrain_gpt.py" --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 1 --global-batch-size 8 --train-iters 100 --lr 1.2e-4 --min-lr 1.2e-5 --lr-decay-style cosine --lr-warmup-fraction 0.01 --weight-decay 0.1 --clip-grad 1.0 \
  --hidden-dropout 0.1 --attention-dropout 0.1 --bf16 --use-distributed-optimizer --overlap-grad-reduce $PAR $DATA $TOK --vocab-size 50257 --log-interval 10 $TINY_ARGS "$@"
