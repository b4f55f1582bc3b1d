#!/bin/bash
# GRPO post-training with packed (THD) rollouts and the phase profiler.
source "$(dirname "$0")/../_common.sh"
ITERS=50; if [ "${TINY:-0}" = "1" ]; then ITERS=2; fi
python "$ROOT/train_rl.py" --preset tiny_llama --train-iters $ITERS --grpo-group-size 4 --grpo-prompts-per-step 4 --grpo-kl-beta 0.01 --grpo-clamp-eps-lower 0.2 --grpo-clamp-eps-upper 0.28 \
  --grpo-filter-groups-with-same-reward --rl-use-sequence-packing --rl-sequence-packing-bin-size 256 --rl-profile --rl-profile-dir "${RL_PROFILE_DIR:-/tmp/rl_profile}" "$@"
