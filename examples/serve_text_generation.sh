#!/bin/bash
# REST server (PUT /api  {"prompts": [...], "tokens_to_generate": 64}) on a checkpoint directory written by pretrain_gpt.py --save.
set -euo pipefail
cd "$(dirname "$0")/.."
CKPT=${1:?usage: serve_text_generation.sh <checkpoint dir> [port]}
python tools/run_text_generation_server.py --load "$CKPT" --preset llama3_8b --engine dynamic --port "${2:-5000}"
