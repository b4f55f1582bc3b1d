#!/bin/bash
# Llama-3 70B on ONE 8xB200 node: TP 8 + SP, full recompute, bf16 main grads (133.9 GiB peak per GPU) — BASELINE config #5.
source "$(dirname "$0")/../_common.sh"
PAR="--tensor-model-parallel-size 8 --sequence-parallel"
if [ "${TINY:-0}" = "1" ]; then PAR=""; fi
$LAUNCH "$ROOT/pretrain_gpt.py" --model llama3_70b --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 4 --train-iters 50 \
  --lr 1.5e-4 --bf16 --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope --untie-embeddings-and-output-weights \
  --recompute-granularity full --recompute-method uniform --recompute-num-layers 1 $PAR $DATA $TOK --vocab-size 128256 --log-interval 5 $TINY_ARGS "$@"
