#!/bin/bash
# Llama-3 8B, sequence 8192, TP=8 + SP on one node.  TP collectives run inside the fused tcgen05 GEMM kernels (MEGATRON_B200_TP_COMM=auto → fused).
set -euo pipefail
cd "$(dirname "$0")/.."
python -m megatron_b200.ops.build
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 pretrain_gpt.py \
  --model llama3_8b --tensor-model-parallel-size 8 --sequence-parallel --bf16 \
  --seq-length 8192 --max-position-embeddings 8192 --micro-batch-size 1 --global-batch-size 16 \
  --train-iters 100 --lr 3e-4 --min-lr 3e-5 --lr-decay-style cosine --lr-warmup-iters 10 --weight-decay 0.1 --clip-grad 1.0 \
  --use-distributed-optimizer --overlap-grad-reduce --overlap-param-gather \
  --mock-data --tokenizer-type NullTokenizer --vocab-size 128255 \
  --log-interval 10 --log-throughput --tensorboard-dir runs/llama3_8b_tp8 "$@"
