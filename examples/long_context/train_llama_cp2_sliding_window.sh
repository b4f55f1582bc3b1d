#!/bin/bash
# Long sequences: context parallel 2 (ring attention on the native block kernels) x tensor parallel 4; add --window-size 4095 0 for a sliding-window model
# (the window runs inside the attention kernels as a band mask).
source "$(dirname "$0")/../_common.sh"
PAR="--tensor-model-parallel-size 4 --context-parallel-size 2 --sequence-parallel --cp-comm-type p2p"
if [ "${TINY:-0}" = "1" ]; then PAR=""; fi
$LAUNCH "$ROOT/pretrain_gpt.py" --model llama3_8b --seq-length 32768 --max-position-embeddings 32768 --micro-batch-size 1 --global-batch-size 4 --train-iters 50 --lr 1e-4 --bf16 \
  --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope --untie-embeddings-and-output-weights --recompute-granularity selective \
  $PAR $DATA $TOK --vocab-size 128256 --log-interval 5 $TINY_ARGS "$@"
