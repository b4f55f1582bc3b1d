#!/bin/bash
# DeepSeek-style stack: multi-latent attention (low-rank q / kv, decoupled rotary key, YaRN) + fine-grained MoE with a shared expert and group-limited routing.
source "$(dirname "$0")/../_common.sh"
MLA="--multi-latent-attention --q-lora-rank 1536 --kv-lora-rank 512 --qk-head-dim 128 --qk-pos-emb-head-dim 64 --v-head-dim 128 --rotary-scaling-factor 40 --mscale 1.0 --mscale-all-dim 1.0"
MOE="--num-experts 32 --moe-router-topk 8 --moe-ffn-hidden-size 2048 --moe-shared-expert-intermediate-size 2048 --moe-grouped-gemm --expert-model-parallel-size 8 --moe-token-dispatcher-type flex"
if [ "${TINY:-0}" = "1" ]; then
  MLA="--multi-latent-attention --q-lora-rank 32 --kv-lora-rank 16 --qk-head-dim 16 --qk-pos-emb-head-dim 8 --v-head-dim 16"
  MOE="--num-experts 4 --moe-router-topk 2 --moe-ffn-hidden-size 64 --moe-shared-expert-intermediate-size 64 --moe-token-dispatcher-type alltoall"
fi
$LAUNCH "$ROOT/pretrain_gpt.py" --num-layers 16 --hidden-size 4096 --num-attention-heads 32 --ffn-hidden-size 11008 --seq-length 4096 --max-position-embeddings 4096 \
  --micro-batch-size 1 --global-batch-size 8 --train-iters 100 --lr 2e-4 --bf16 --swiglu --normalization RMSNorm --disable-bias-linear --position-embedding-type rope \
  --untie-embeddings-and-output-weights $MLA $MOE $DATA $TOK --vocab-size 102400 --log-interval 10 $TINY_ARGS "$@"
