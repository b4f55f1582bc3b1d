#!/usr/bin/env python
"""Knowledge distillation (teacher logits → student, KL on the vocabulary-parallel logits) followed by post-training weight quantisation of the student with a
per-layer recipe.

    python examples/post_training/distill_and_quantize.py --iters 5
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29579")
    dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model
    from megatron_b200.post_training.distillation import DistillationModel

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    cpu = not torch.cuda.is_available()
    teacher, _, p = build_gpt_model("tiny_llama", use_cpu_initialization=cpu)
    student, _, _ = build_gpt_model("tiny_llama", use_cpu_initialization=cpu, num_layers=1)
    dm = DistillationModel(student, teacher, alpha=0.5, temperature=2.0)
    opt = torch.optim.AdamW(student.parameters(), lr=1e-3)
    dev = next(student.parameters()).device
    for it in range(args.iters):
        tokens = torch.randint(0, p["vocab_size"], (2, 32), device=dev)
        pos = torch.arange(32, device=dev)[None].expand(2, -1)
        loss, parts = dm(tokens, pos, None, labels=tokens.roll(-1, 1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        print(f"iter {it + 1}: loss {float(loss):.4f}  (ce {float(parts['ce']):.4f}, kd {float(parts['kd']):.4f})", flush=True)
    from megatron_b200.post_training.quantize import PTQConfig, quantize_model

    states = quantize_model(student, PTQConfig(default="fp8", matchers=[("*output_layer*", "none")]))
    with torch.no_grad():
        tokens = torch.randint(0, p["vocab_size"], (1, 16), device=dev)
        out = student(tokens, torch.arange(16, device=dev)[None], None)
    print(f"quantised {len(states)} linear layers; forward after quantisation: logits {tuple(out.shape)}, finite = {bool(torch.isfinite(out).all())}")


if __name__ == "__main__":
    main()
