#!/bin/bash
# Llama-3 8B, TP 8 + sequence parallel on the fused AG->GEMM / GEMM->RS kernels, distributed optimizer over NVLink — the benchmark configuration.
source "$(dirname "$0")/../_common.sh"
PAR="--tensor-model-parallel-size 8 --sequence-parallel"
if [ "${TINY:-0}" = "1" ]; then PAR=""; fi
$LAUNCH "$ROOT/pretrain_gpt.py" --model llama3_8b --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 16 --train-iters 100 \
  --lr 3e-4 --min-lr 3e-5 --lr-decay-style cosine --weight-decay 0.1 --clip-grad 1.0 --bf16 --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope \
  --untie-embeddings-and-output-weights --use-distributed-optimizer --recompute-granularity selective $PAR $DATA $TOK --vocab-size 128256 --log-interval 10 $TINY_ARGS "$@"
