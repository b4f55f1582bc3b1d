# Shared launcher for the example scripts.  Environment:
#   NPROC (default 8)   ranks on this node                MASTER_PORT (default 29500)
#   TINY=1              shrink the model / run a few iterations — CPU-friendly smoke mode (gloo, used by tests/test_examples_cpu.py)
#   DATA_PATH           a preprocessed .bin/.idx prefix; without it the scripts train on synthetic tokens (--mock-data)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NPROC=${NPROC:-8}
MASTER_PORT=${MASTER_PORT:-29500}
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NPROC} --master-addr 127.0.0.1 --master-port ${MASTER_PORT}"
if [ -n "${DATA_PATH:-}" ]; then DATA="--data-path ${DATA_PATH} --split 949,50,1"; else DATA="--mock-data"; fi
TOK="--tokenizer-type NullTokenizer"
if [ "${TINY:-0}" = "1" ]; then
  export CUDA_VISIBLE_DEVICES=""
  # every size flag below overrides what the script passed before it (argparse: the last occurrence wins)
  TINY_ARGS="--num-layers 2 --hidden-size 64 --num-attention-heads 4 --num-query-groups 4 --ffn-hidden-size 128 --seq-length 64 --max-position-embeddings 64 --vocab-size 1024 \
    --micro-batch-size 1 --global-batch-size ${NPROC} --train-iters 2 --log-interval 1 --eval-iters 0 --distributed-backend gloo --lr-decay-iters 100 --no-bf16-tiny"
else
  TINY_ARGS=""
fi
