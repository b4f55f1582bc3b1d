#!/usr/bin/env python
"""Headline benchmark: Llama-3-8B training throughput, TP=N (+sequence parallel), bf16.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 5 --warmup 3
    python bench.py --impl reference ...      # unmodified NVIDIA/Megatron-LM from baseline/_ref

Prints ONE JSON line on rank 0 (contract in the task statement).  `value` = whole-job tokens/s,
device-timed with CUDA events around exactly K optimizer steps, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama3_8b")
    ap.add_argument("--global-batch", type=int, default=4, help="sequences per optimizer step")
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None, help="DEV ONLY: fewer layers (result is flagged invalid)")
    ap.add_argument("--tp-comm", default=None, choices=[None, "nccl", "nvlink", "fused"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--recompute", default="auto")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def timed_loop(torch, dist, step_fn, steps, world):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns max-over-ranks ms."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1), wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0]), float(ms[1])


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, REPO)
    rank, world, local = env_rank()
    assert world == args.gpus, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    if args.tp_comm:
        os.environ["MEGATRON_B200_TP_COMM"] = args.tp_comm
    from megatron_b200 import ops
    from megatron_b200.training.engine import TrainEngine

    torch.cuda.set_device(local)
    overrides = {}
    if args.layers:
        overrides["num_layers"] = args.layers
    recompute = {}
    if args.recompute == "auto":
        # 1 GPU holds all 8B parameters + optimizer state: recompute the cheap ops' outputs
        recompute = dict(recompute_granularity="selective", recompute_modules=["layernorm", "mlp_act"]) if world == 1 else {}
    elif args.recompute == "full":
        recompute = dict(recompute_granularity="full", recompute_method="uniform", recompute_num_layers=1)
    eng = TrainEngine(args.model, tensor_model_parallel_size=world, sequence_parallel=world > 1, micro_batch_size=args.micro_batch,
                      global_batch_size=args.global_batch, seq_length=args.seq, bf16=True, model_overrides=overrides, **recompute)
    host = eng.synthetic_batch(pinned=True, seed=rank * 0 + 17)
    dev_tokens = host.to("cuda")
    last = {}

    def step_dev():
        last["loss"] = eng.train_step(dev_tokens)

    def step_e2e():
        last["loss_host"] = float(eng.train_step(host))  # H2D of inputs + D2H of the loss every step

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    ms, wall_ms = timed_loop(torch, dist, step_dev, args.steps, world)
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else {}
    tokens_per_step = args.global_batch * eng.seq_length
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        step_e2e()
        _, e_wall = timed_loop(torch, dist, step_e2e, args.steps, world)
        e2e = {"value": tokens_per_step * args.steps / (e_wall / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": host.numel() * host.element_size(),
               "d2h_bytes_per_step": 4, "timing": "host wall clock incl. H2D inputs from pinned memory and D2H loss read, max over ranks"}
    peak = torch.cuda.max_memory_allocated() / 2**30
    from megatron_b200.ops import gemm as _gemm
    from megatron_b200.parallel import fused as _fusedmod

    _tp_mode = _fusedmod.get_mode(world_size=world) if world > 1 else "n/a"
    _gemm_wins = {}
    for _key, _winner, _times in _gemm.tuning_report():
        _gemm_wins[str(_winner).split(":")[0]] = _gemm_wins.get(str(_winner).split(":")[0], 0) + 1
    _attn_impl = ops.attention_impl_for(eng.seq_length, args.micro_batch, 32 // world) if hasattr(ops, "attention_impl_for") else "library"
    if rank == 0:
        flops = eng.flops_per_step * args.steps / (ms / 1e3)
        mp = {}
        try:
            mp = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = mp.get("bf16_tflops_sustained", 1400.0)
        out = {
            "metric": "tokens/sec (whole job, device-timed, max over ranks), Llama-3-8B TP=N + sequence parallel", "value": value, "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random token ids of the named shape; random-init weights)",
            "impl": "b200", "tflops_per_gpu": flops / world / 1e12, "mfu_of_measured_sustained_cublas": flops / world / 1e12 / peak_tf,
            "loss": float(last["loss"]), "peak_mem_gib": peak, "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
            "config": {"model": args.model + (f"[layers={args.layers} DEV-INVALID]" if args.layers else ""), "global_batch": args.global_batch,
                       "micro_batch": args.micro_batch, "seq_len": eng.seq_length, "parallelism": f"tp{world}" + ("+sp" if world > 1 else ""),
                       "l2_policy": "inputs larger than L2 (16 GB of bf16 weights + activations stream through the 126 MB L2 every step)",
                       "main_grads": "bf16", "optimizer": "fused AdamW, fp32 master+moments", "recompute": recompute or "none",
                       "tp_comm": os.environ.get("MEGATRON_B200_TP_COMM", "auto"), "tp_comm_resolved": _tp_mode, "gemm": os.environ.get("MEGATRON_B200_GEMM", "auto"),
                       "plain_gemm_shapes_tuned": _gemm_wins, "attention": _attn_impl},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The UNMODIFIED reference (baseline/_ref, megatron-core 0.20.0) through its public API:
    parallel_state → GPTModel(local spec) → DistributedDataParallel → get_megatron_optimizer →
    get_forward_backward_func → finalize_model_grads.  TransformerEngine/Apex are not installed, so
    this is the reference's stock "local" path (cuBLAS + NCCL, unfused attention, torch AdamW).
    The local backend asserts `not sequence_parallel` for its norm layers, so TP=N runs WITHOUT
    sequence parallelism (all-reduce TP) — the closest configuration the stock code supports."""
    ref = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "megatron", "core")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed (pip install --no-deps --target baseline/_ref /root/reference)"}))
        return
    sys.path.insert(0, ref)
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    import warnings

    warnings.filterwarnings("ignore")
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank()
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    try:
        from functools import partial

        import torch.nn.functional as F
        from megatron.core import parallel_state
        from megatron.core.distributed import DistributedDataParallel, DistributedDataParallelConfig
        from megatron.core.distributed.finalize_model_grads import finalize_model_grads
        from megatron.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
        from megatron.core.models.gpt.gpt_model import GPTModel
        from megatron.core.optimizer import OptimizerConfig, get_megatron_optimizer
        from megatron.core.pipeline_parallel.schedules import get_forward_backward_func
        from megatron.core.tensor_parallel.random import model_parallel_cuda_manual_seed
        from megatron.core.transformer.transformer_config import TransformerConfig
    except Exception as e:  # pragma: no cover
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"import failed: {type(e).__name__}: {e}"[:300]}))
        return

    P = dict(num_layers=args.layers or 32, hidden=4096, ffn=14336, heads=32, groups=8, kv=128, vocab=128256, seq=args.seq or 8192)
    assert args.model == "llama3_8b", "reference arm implements the headline config only"
    parallel_state.initialize_model_parallel(tensor_model_parallel_size=world)
    model_parallel_cuda_manual_seed(1234)

    def build(recompute):
        cfg = TransformerConfig(
            num_layers=P["num_layers"], hidden_size=P["hidden"], ffn_hidden_size=P["ffn"], num_attention_heads=P["heads"], num_query_groups=P["groups"],
            kv_channels=P["kv"], normalization="RMSNorm", gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False, hidden_dropout=0.0,
            attention_dropout=0.0, tensor_model_parallel_size=world, sequence_parallel=False, bf16=True, params_dtype=torch.bfloat16,
            bias_activation_fusion=True, bias_dropout_fusion=True, masked_softmax_fusion=True, gradient_accumulation_fusion=False, **recompute,
        )
        m = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=P["vocab"], max_sequence_length=P["seq"], parallel_output=True,
                     share_embeddings_and_output_weights=False, position_embedding_type="rope", rotary_base=500000).cuda()
        ddp = DistributedDataParallel(cfg, DistributedDataParallelConfig(grad_reduce_in_fp32=False, overlap_grad_reduce=False, use_distributed_optimizer=True), m)
        opt = get_megatron_optimizer(OptimizerConfig(optimizer="adam", lr=3e-4, min_lr=3e-5, weight_decay=0.1, bf16=True, params_dtype=torch.bfloat16,
                                                     clip_grad=1.0, use_distributed_optimizer=True, adam_beta1=0.9, adam_beta2=0.95), [ddp])
        cfg.finalize_model_grads_func = finalize_model_grads
        cfg.no_sync_func = ddp.no_sync
        return cfg, ddp, opt

    nmb = args.global_batch // args.micro_batch
    g = torch.Generator().manual_seed(17)
    host = torch.randint(0, P["vocab"], (nmb * args.micro_batch, P["seq"] + 1), generator=g, dtype=torch.int64).pin_memory()
    dev_tokens = host.cuda()
    pos = torch.arange(P["seq"], device="cuda").unsqueeze(0).expand(args.micro_batch, -1).contiguous()
    fwd_bwd = get_forward_backward_func()

    def loss_func(out):
        loss = out.float().mean()
        return loss, {"lm loss": loss.detach()}

    def fstep(it, model):
        b = next(it)
        return model(b[:, :-1].contiguous(), pos, None, labels=b[:, 1:].contiguous()), loss_func

    state = {}

    def make_step(tokens_fn, sync_loss):
        def step():
            t = tokens_fn()
            state["ddp"].zero_grad_buffer()
            state["opt"].zero_grad()
            it = iter(t[i * args.micro_batch : (i + 1) * args.micro_batch] for i in range(nmb))
            losses = fwd_bwd(forward_step_func=fstep, data_iterator=it, model=state["ddp"], num_microbatches=nmb, seq_length=P["seq"],
                             micro_batch_size=args.micro_batch, forward_only=False)
            state["opt"].step()
            l = torch.stack([d["lm loss"] for d in losses]).mean()
            state["loss"] = float(l) if sync_loss else l
        return step

    full = dict(recompute_granularity="full", recompute_method="uniform", recompute_num_layers=1)
    attempts = [("selective(core_attn)", dict(recompute_granularity="selective", recompute_modules=["core_attn"])),
                ("full(uniform,1)", full),
                ("full(uniform,1)+bf16-softmax", dict(full, attention_softmax_in_fp32=False))]
    if world == 1:
        # measured on the 180 GB part: selective recompute OOMs at TP=1 (fp32 [32,8192,8192] scores on top of
        # 120 GiB of parameter/optimizer state) and a failed attempt fragments the heap — start from full recompute
        attempts = attempts[1:]
    if os.environ.get("REF_RECOMPUTE"):
        attempts = [a for a in attempts if a[0].startswith(os.environ["REF_RECOMPUTE"])] or attempts
    used = None
    for name, rc in attempts:
        try:
            state.clear()
            torch.cuda.empty_cache()
            state["cfg"], state["ddp"], state["opt"] = build(rc)
            step_dev = make_step(lambda: dev_tokens, False)
            for _ in range(args.warmup):
                step_dev()
            torch.cuda.synchronize()
            used = name
            break
        except torch.cuda.OutOfMemoryError as e:
            if rank == 0:
                print(f"[reference] OOM with recompute={name}: {str(e)[:300]}", file=sys.stderr, flush=True)
            state.clear()
            import gc

            gc.collect()
            torch.cuda.empty_cache()
            continue
    ok = torch.tensor([1 if used else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "reference local path runs out of memory on this config even with full activation recompute"}))
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _ = timed_loop(torch, dist, step_dev, args.steps, world)
    clocks = sampler.stop() if rank == 0 else {}
    tokens_per_step = args.global_batch * P["seq"]
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        step_e2e = make_step(lambda: host.cuda(non_blocking=True), True)
        step_e2e()
        _, e_wall = timed_loop(torch, dist, step_e2e, args.steps, world)
        e2e = {"value": tokens_per_step * args.steps / (e_wall / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": host.numel() * 8, "d2h_bytes_per_step": 4}
    if rank == 0:
        loss = state.get("loss")
        print(json.dumps({
            "impl": "reference", "metric": "tokens/sec (whole job, device-timed, max over ranks), Llama-3-8B TP=N", "value": value, "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "loss": float(loss) if loss is not None else None, "clocks": clocks, "e2e": e2e,
            "peak_mem_gib": torch.cuda.max_memory_allocated() / 2**30,
            "config": {"model": args.model + (f"[layers={args.layers} DEV-INVALID]" if args.layers else ""), "global_batch": args.global_batch,
                       "micro_batch": args.micro_batch, "seq_len": P["seq"], "parallelism": f"tp{world} (no SP: local backend asserts not sequence_parallel)",
                       "recompute": used, "backend": "megatron-core 0.20.0 local spec (TE/Apex absent): cuBLAS + NCCL + torch AdamW, unfused attention"},
        }), flush=True)
    if world > 1:
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
