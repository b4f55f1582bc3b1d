"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities) for a reduced-depth Llama-3 8B (same layer shapes).

    python tools/step_profile.py [--layers 4] [--tp-comm nccl]      (single GPU;  under torchrun for TP > 1)
Prints the top kernels by summed device time and their share; numbers are for attribution only (profiler overhead), never a bench value."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--tp-comm", default=None)
    ap.add_argument("--top", type=int, default=28)
    args = ap.parse_args()
    if args.tp_comm:
        os.environ["MEGATRON_B200_TP_COMM"] = args.tp_comm
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from megatron_b200.training.engine import TrainEngine

    eng = TrainEngine("llama3_8b", tensor_model_parallel_size=world, sequence_parallel=world > 1, micro_batch_size=1, global_batch_size=4, bf16=True,
                      model_overrides={"num_layers": args.layers},
                      **(dict(recompute_granularity="selective", recompute_modules=["layernorm", "mlp_act"]) if world == 1 else {}))
    batch = eng.synthetic_batch().cuda()
    for _ in range(3):
        eng.train_step(batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.train_step(batch)
        torch.cuda.synchronize()
    if rank == 0:
        ev = [e for e in prof.key_averages() if e.device_time_total > 0]
        tot = sum(e.device_time_total for e in ev)
        print(f"layers={args.layers} world={world}: total device time {tot / 1e3:.1f} ms in {sum(e.count for e in ev)} kernels")
        for e in sorted(ev, key=lambda e: -e.device_time_total)[: args.top]:
            print(f"{100 * e.device_time_total / tot:5.1f}%  {e.device_time_total / 1e3:8.2f} ms  x{e.count:<5d} {e.key[:120]}")


if __name__ == "__main__":
    main()
