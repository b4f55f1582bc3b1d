#!/bin/bash
# REST / OpenAI-compatible text-generation server on the continuous-batching engine: paged KV cache, paged-attention decode kernel, CUDA-graphed decode buckets,
# chunked prefill (2048 prompt tokens per step), prefix caching.
source "$(dirname "$0")/../_common.sh"
python "$ROOT/tools/run_text_generation_server.py" --preset "${PRESET:-tiny_llama}" --port "${PORT:-5000}" --inference-dynamic-batching \
  --inference-dynamic-batching-block-size 16 --inference-dynamic-batching-max-requests 64 --enable-chunked-prefill --inference-dynamic-batching-prefix-caching \
  --inference-dynamic-batching-num-cuda-graphs 8 "$@"
