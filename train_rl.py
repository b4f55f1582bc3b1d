#!/usr/bin/env python
"""GRPO post-training of a GPT preset on a verifiable toy task (drop-in for the reference's ``train_rl.py`` entry point).

    python train_rl.py --preset tiny_llama --iters 20 --group-size 8
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny_llama")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--group-size", type=int, default=4)
    ap.add_argument("--prompts-per-iter", type=int, default=4)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--vocab", type=int, default=64)
    # the reference's flag names (megatron/training/arguments.py _add_rl_args); they take precedence over the short spellings above
    ap.add_argument("--grpo-group-size", type=int, default=None)
    ap.add_argument("--grpo-prompts-per-step", type=int, default=None)
    ap.add_argument("--grpo-iterations", type=int, default=1, help="optimisation passes over every rollout batch")
    ap.add_argument("--grpo-kl-beta", type=float, default=None)
    ap.add_argument("--grpo-clamp-eps-lower", type=float, default=None)
    ap.add_argument("--grpo-clamp-eps-upper", type=float, default=None)
    ap.add_argument("--grpo-entropy-term-weight", type=float, default=None)
    ap.add_argument("--grpo-filter-groups-with-same-reward", action="store_true")
    ap.add_argument("--rl-default-temperature", type=float, default=None)
    ap.add_argument("--perform-rl-step", action="store_true", help="accepted: this entry point always performs RL steps")
    ap.add_argument("--train-iters", type=int, default=None)
    ap.add_argument("--rl-use-sequence-packing", action="store_true", help="log-probs / training on packed (THD) rows: no padding enters the model")
    ap.add_argument("--rl-sequence-packing-bin-size", type=int, default=512)
    ap.add_argument("--rl-profile", action="store_true", help="time the rollout / log-prob / train phases of every iteration")
    ap.add_argument("--rl-profile-dir", default=None)
    args = ap.parse_args()
    if args.grpo_group_size:
        args.group_size = args.grpo_group_size
    if args.grpo_prompts_per_step:
        args.prompts_per_iter = args.grpo_prompts_per_step
    if args.train_iters:
        args.iters = args.train_iters
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29571")
    if not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model
    from megatron_b200.rl.grpo import CountTokenEnv, GRPOConfig, GRPOTrainer

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1234)
    model, _, p = build_gpt_model(args.preset, use_cpu_initialization=not torch.cuda.is_available())
    ref, _, _ = build_gpt_model(args.preset, use_cpu_initialization=not torch.cuda.is_available())  # frozen reference policy
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.AdamW(model.parameters(), lr=args.lr)
    cfg = GRPOConfig(group_size=args.group_size, max_new_tokens=8)
    if args.grpo_kl_beta is not None:
        cfg.kl_beta = args.grpo_kl_beta
    if args.grpo_clamp_eps_lower is not None:
        cfg.clip_eps = args.grpo_clamp_eps_lower
    if args.grpo_clamp_eps_upper is not None:
        cfg.clip_eps_upper = args.grpo_clamp_eps_upper
    if args.grpo_entropy_term_weight is not None:
        cfg.entropy_coef = args.grpo_entropy_term_weight
    if args.rl_default_temperature is not None:
        cfg.temperature = args.rl_default_temperature
    cfg.filter_groups_with_same_reward = args.grpo_filter_groups_with_same_reward
    cfg.use_sequence_packing, cfg.packing_bin_size = args.rl_use_sequence_packing, args.rl_sequence_packing_bin_size
    tr = GRPOTrainer(model, ref, opt, CountTokenEnv(args.vocab), cfg, vocab_size=args.vocab)
    if args.rl_profile:
        from megatron_b200.rl.rl_profiling import RLProfiler

        tr.profiler = RLProfiler(enabled=True, out_dir=args.rl_profile_dir or ".")
    for it in range(args.iters):
        s = tr.step(args.prompts_per_iter, inner_epochs=args.grpo_iterations)
        print(f"iter {it + 1:3d} | reward {float(s['reward']):.3f} | loss {float(s['loss']):+.4f} | kl {float(s['kl']):.5f}", flush=True)
    if args.rl_profile:
        path = tr.profiler.dump()
        print(f"rl profile: {path}  summary: {tr.profiler.summary()}", flush=True)
    return tr


if __name__ == "__main__":
    main()
