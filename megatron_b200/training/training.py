"""``pretrain()`` — argument-driven training loop (reference ``megatron/training/training.py:1500-4554``).

init distributed + model-parallel groups → build model chunks (per virtual stage) → DDP → optimizer +
LR scheduler → (load checkpoint) → data iterators → ``train()``: step / log / eval / checkpoint / exit.
"""
from __future__ import annotations

import gc
import math
import os
import signal
import sys
import time
from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from ..core import parallel_state as ps
from ..core.distributed import DistributedDataParallel, DistributedDataParallelConfig, finalize_model_grads
from ..core.enums import ModelType
from ..core.num_microbatches_calculator import (
    destroy_num_microbatches_calculator,
    get_current_global_batch_size,
    get_num_microbatches,
    init_num_microbatches_calculator,
    update_num_microbatches,
)
from ..core.optimizer import OptimizerConfig, get_megatron_optimizer
from ..core.optimizer_param_scheduler import OptimizerParamScheduler
from ..core.pipeline_parallel.schedules import get_forward_backward_func
from ..core.rerun_state_machine import RerunDataIterator, get_rerun_state_machine, initialize_rerun_state_machine
from ..core.tensor_parallel.random import model_parallel_cuda_manual_seed
from ..core.timers import Timers
from ..core.utils import StragglerDetector
from . import checkpointing
from .arguments import parse_and_validate_args
from .engine import initialize_distributed
from .flops import num_floating_point_operations

_GLOBALS: Dict[str, object] = {}


def get_args():
    return _GLOBALS["args"]


def get_timers() -> Timers:
    return _GLOBALS["timers"]


def print_rank_0(*a):
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(*a, flush=True)


def print_rank_last(*a):
    if not dist.is_initialized() or dist.get_rank() == dist.get_world_size() - 1:
        print(*a, flush=True)


# ---------------------------------------------------------------------------------------------------------
def initialize_megatron(argv=None, extra_args_provider=None, args_defaults: Optional[dict] = None):
    args = parse_and_validate_args(argv, extra_args_provider=extra_args_provider)
    for k, v in (args_defaults or {}).items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)
    _GLOBALS["args"] = args
    if getattr(args, "nccl_flight_recorder_dir", None):
        from .initialize import setup_nccl_flight_recorder

        setup_nccl_flight_recorder(args.nccl_flight_recorder_dir)
    initialize_distributed("gloo" if (args.distributed_backend == "gloo" or not torch.cuda.is_available()) else "nccl")
    if not ps.is_initialized():
        ps.initialize_model_parallel(
            tensor_model_parallel_size=args.tensor_model_parallel_size, pipeline_model_parallel_size=args.pipeline_model_parallel_size,
            virtual_pipeline_model_parallel_size=args.virtual_pipeline_model_parallel_size, context_parallel_size=args.context_parallel_size,
            expert_model_parallel_size=args.expert_model_parallel_size, expert_tensor_parallel_size=args.expert_tensor_parallel_size,
            distributed_timeout_minutes=args.distributed_timeout_minutes, create_gloo_process_groups=False,
            order="tp-cp-ep-pp-dp" if getattr(args, "use_tp_pp_dp_mapping", False) else "tp-cp-ep-dp-pp",
            nccl_communicator_config_path=getattr(args, "nccl_communicator_config_path", None), high_priority_stream_groups=getattr(args, "high_priority_stream_groups", None) or None,
        )
    # --data-parallel-random-init: every data-parallel replica draws from its own stream (parameters are broadcast from the first replica after wrapping)
    model_parallel_cuda_manual_seed(args.seed + (10 * ps.get_data_parallel_rank() if getattr(args, "data_parallel_random_init", False) else 0))
    if getattr(args, "batch_invariant_mode", False):
        from ..core.transformer.custom_layers.batch_invariant_kernels import enable_batch_invariant_mode

        enable_batch_invariant_mode()
    if getattr(args, "enable_experimental", False):
        from ..core import config as _core_config

        _core_config.set_experimental_flag(True)
    if getattr(args, "logging_level", None) is not None:
        import logging

        logging.getLogger().setLevel(args.logging_level)
    _GLOBALS["timers"] = Timers(args.timing_log_level, args.timing_log_option)
    from . import global_vars

    args.world_size = dist.get_world_size()
    global_vars.set_global_variables(args)
    _GLOBALS["tensorboard"], _GLOBALS["wandb"], _GLOBALS["one_logger"] = (global_vars.get_tensorboard_writer(), global_vars.get_wandb_writer(),
                                                                          global_vars.get_one_logger())
    if args.deterministic_mode:
        from .determinism import enable_deterministic_mode

        enable_deterministic_mode(strict=False)
    if getattr(args, "enable_ft_package", False):
        from . import ft_integration

        ft_integration.setup(args, dist.get_rank())
        ft_integration.maybe_setup_simulated_fault(args, dist.get_rank(), dist.get_world_size())
    initialize_rerun_state_machine(mode=args.rerun_mode, error_injection_rate=args.error_injection_rate, error_injection_type=getattr(args, "error_injection_type", "transient_error"))
    destroy_num_microbatches_calculator()
    schedule = getattr(args, "step_batch_size_schedule", None)
    init_num_microbatches_calculator(dist.get_rank(), args.rampup_batch_size, args.global_batch_size, args.micro_batch_size, args.data_parallel_size,
                                     decrease_batch_size_if_needed=bool(getattr(args, "decrease_batch_size_if_needed", False)), step_batch_size_schedule=schedule,
                                     seq_length=args.seq_length if schedule else None)
    update_train_iters(args)
    from ..core.fault_injector import FaultInjector, FaultInjectorConfig

    fi_cfg = FaultInjectorConfig.from_args(args, dist.get_world_size())
    _GLOBALS["fault_injector"] = FaultInjector(fi_cfg) if fi_cfg is not None else None
    if args.tp_comm != "auto":
        from ..parallel import fused

        fused.set_mode(args.tp_comm)
    if torch.cuda.is_available() and args.tensor_model_parallel_size > 1 and dist.get_backend() == "nccl":
        from ..parallel import collectives, fused

        if fused.get_mode(world_size=args.tensor_model_parallel_size) != "nccl":
            try:
                collectives.enable_for_group(ps.get_tensor_model_parallel_group())
            except Exception as e:  # no symmetric memory on this box: NCCL + GEMM path
                print_rank_0(f"WARNING: NVLink symmetric-memory runtime unavailable ({type(e).__name__}: {e}); TP collectives use NCCL")
                fused.set_mode("nccl")
        if args.expert_model_parallel_size > 1 and getattr(args, "moe_token_dispatcher_type", None) == "flex":
            collectives.enable_for_group(ps.get_expert_model_parallel_group())
    return args


def get_model(model_provider_func: Callable, wrap_with_ddp: bool = True) -> List[torch.nn.Module]:
    """One model chunk per virtual pipeline stage, moved to the device and wrapped in DDP (reference :2358)."""
    args = get_args()
    vp = args.virtual_pipeline_model_parallel_size
    chunks = []
    for v in range(vp or 1):
        if vp:
            ps.set_virtual_pipeline_model_parallel_rank(v)
        pre = ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=v if vp else None)
        post = ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=v if vp else None)
        m = model_provider_func(pre_process=pre, post_process=post, vp_stage=v if vp else None)
        m.model_type = ModelType.encoder_or_decoder
        chunks.append(m)
    if vp:
        ps.set_virtual_pipeline_model_parallel_rank(0)
    n_params = sum(p.numel() for c in chunks for p in c.parameters())
    if ps.get_data_parallel_rank() == 0:
        print(f" > number of parameters on (tensor, pipeline) model parallel rank ({ps.get_tensor_model_parallel_rank()}, "
              f"{ps.get_pipeline_model_parallel_rank()}): {n_params}", flush=True)
    dev = torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and dist.get_backend() != "gloo") else torch.device("cpu")
    chunks = [c.to(dev) for c in chunks]
    if not wrap_with_ddp:
        return chunks
    config = chunks[0].config
    from .arguments import ddp_config_from_args

    if getattr(args, "use_megatron_fsdp", False) or getattr(args, "use_torch_fsdp2", False):
        # parameters / gradients / optimizer state sharded over the data-parallel group (reference --use-megatron-fsdp + --data-parallel-sharding-strategy,
        # --use-torch-fsdp2); this build supports it for TP = PP = 1
        assert args.tensor_model_parallel_size == 1 and args.pipeline_model_parallel_size == 1, "FSDP in the training loop needs TP = PP = 1 in this build"
        ddp_config = ddp_config_from_args(args)
        if getattr(args, "use_torch_fsdp2", False):
            from ..core.distributed.torch_fully_sharded_data_parallel import TorchFullyShardedDataParallel

            model = [TorchFullyShardedDataParallel(config, ddp_config, c, reshard_after_forward=getattr(args, "torch_fsdp2_reshard_after_forward", True)) for c in chunks]
        else:
            from ..core.distributed.fsdp.mcore_fsdp_adapter import wrap_model_with_fsdp

            model = wrap_model_with_fsdp(chunks, config, ddp_config, strategy=getattr(args, "data_parallel_sharding_strategy", None) or "optim_grads_params")
            if ps.get_data_parallel_world_size() > 1:
                for m in model:
                    m.broadcast_params()
        for c in chunks:
            c.config = config
        return model

    ddp_config = ddp_config_from_args(args)
    if getattr(args, "ddp_num_buckets", None):
        n_params = sum(p.numel() for c in chunks for p in c.parameters())
        ddp_config.bucket_size = max(1, -(-n_params // args.ddp_num_buckets))
    model = [DistributedDataParallel(config, ddp_config, c, disable_bucketing=(i > 0)) for i, c in enumerate(chunks)]
    for c in chunks:
        c.config = config
    if ps.get_data_parallel_world_size() > 1:
        for m in model:
            m.broadcast_params()
    return model


def update_train_iters(args) -> None:
    """Sample-based runs (``--train-samples``): the iteration count follows the batch-size schedule (reference ``training.update_train_iters`` :2155)."""
    if not getattr(args, "train_samples", None) or (args.train_iters and not getattr(args, "_train_iters_from_samples", False)):
        return
    if getattr(args, "step_batch_size_schedule", None) is not None or args.rampup_batch_size:
        iters, consumed = 0, 0
        while consumed < args.train_samples:
            update_num_microbatches(consumed, consistency_check=False)
            consumed += get_current_global_batch_size()
            iters += 1
        update_num_microbatches(0, consistency_check=False)
        args.train_iters = iters
    else:
        args.train_iters = args.train_samples // args.global_batch_size
    print_rank_0(f"setting training iterations to {args.train_iters}")


def consumed_samples_at(iteration: int) -> int:
    """Samples consumed after ``iteration`` steps under the active batch-size schedule (constant: ``iteration * global_batch_size``)."""
    consumed = 0
    for _ in range(iteration):
        update_num_microbatches(consumed, consistency_check=False)
        consumed += get_current_global_batch_size()
    update_num_microbatches(consumed, consistency_check=False)
    return consumed


def get_optimizer_param_scheduler(optimizer):
    args = get_args()
    gbs = args.global_batch_size
    if args.train_iters and not getattr(args, "train_samples", None):
        decay = (args.lr_decay_iters or args.train_iters) * gbs
        wd_steps = args.train_iters * gbs
        warm = args.lr_warmup_fraction * decay if args.lr_warmup_fraction is not None else args.lr_warmup_iters * gbs
        wsd = args.lr_wsd_decay_iters * gbs if args.lr_wsd_decay_iters else None
    else:
        decay = args.lr_decay_samples or args.train_samples
        wd_steps = args.train_samples
        warm = args.lr_warmup_fraction * decay if args.lr_warmup_fraction is not None else args.lr_warmup_samples
        wsd = getattr(args, "lr_wsd_decay_samples", None)
    return OptimizerParamScheduler(
        optimizer, init_lr=args.lr_warmup_init, max_lr=args.lr, min_lr=args.min_lr, lr_warmup_steps=int(warm), lr_decay_steps=max(int(decay), int(warm) + 1),
        lr_decay_style=args.lr_decay_style, start_wd=args.start_weight_decay if args.start_weight_decay is not None else args.weight_decay,
        end_wd=args.end_weight_decay if args.end_weight_decay is not None else args.weight_decay, wd_incr_steps=max(int(wd_steps), 1),
        wd_incr_style=args.weight_decay_incr_style, use_checkpoint_opt_param_scheduler=bool(getattr(args, "use_checkpoint_opt_param_scheduler", False)),
        override_opt_param_scheduler=bool(getattr(args, "override_opt_param_scheduler", False)), wsd_decay_steps=wsd, lr_wsd_decay_style=args.lr_wsd_decay_style,
    )


def setup_model_and_optimizer(model_provider_func: Callable):
    args = get_args()
    model = get_model(model_provider_func)
    opt_cfg = OptimizerConfig(
        optimizer=args.optimizer, lr=args.lr, min_lr=args.min_lr, weight_decay=args.weight_decay, fp16=args.fp16, bf16=args.bf16, params_dtype=args.params_dtype,
        clip_grad=args.clip_grad, use_distributed_optimizer=args.use_distributed_optimizer, adam_beta1=args.adam_beta1, adam_beta2=args.adam_beta2,
        adam_eps=args.adam_eps, loss_scale=args.loss_scale, initial_loss_scale=args.initial_loss_scale, min_loss_scale=args.min_loss_scale,
        loss_scale_window=args.loss_scale_window, hysteresis=args.hysteresis, overlap_param_gather=args.overlap_param_gather, timers=get_timers(),
        muon_momentum=getattr(args, "muon_momentum", 0.95), muon_ns_steps=getattr(args, "muon_ns_steps", 5), muon_tp_mode=getattr(args, "muon_tp_mode", "blockwise"),
    )
    if getattr(args, "use_megatron_fsdp", False):
        from ..core.distributed.fsdp.mcore_fsdp_adapter import FSDPOptimizer

        optimizer = FSDPOptimizer(model, opt_cfg, ps.get_model_parallel_group())
    else:
        optimizer = get_megatron_optimizer(opt_cfg, model)
    scheduler = get_optimizer_param_scheduler(optimizer)
    args.iteration, args.num_floating_point_operations_so_far = 0, 0.0
    if args.load or getattr(args, "pretrained_checkpoint", None):
        it, fl, src = checkpointing.load_latest_checkpoint(
            model, optimizer, scheduler, args.load, getattr(args, "non_persistent_global_ckpt_dir", None), getattr(args, "non_persistent_local_ckpt_dir", None),
            getattr(args, "pretrained_checkpoint", None), getattr(args, "ckpt_step", None), getattr(args, "exit_on_missing_checkpoint", False),
            load_optim=not (args.no_load_optim or args.finetune), load_rng=not (args.no_load_rng or args.finetune),
            fully_parallel_load=getattr(args, "ckpt_fully_parallel_load", False), dist_ckpt_strictness=getattr(args, "dist_ckpt_strictness", None))
        args.iteration = 0 if args.finetune else it
        args.num_floating_point_operations_so_far = fl
        args.consumed_train_samples = consumed_samples_at(args.iteration)
        print_rank_0(f" > loaded {src} checkpoint (from {args.load}) at iteration {it}")
    return model, optimizer, scheduler


def append_to_progress_log(string: str, barrier: bool = True) -> None:
    """``--log-progress``: one line per job start / checkpoint / exit in ``<save>/progress.txt`` with the job id, world size and cumulative FLOPs — enough to
    reconstruct a run's throughput history across restarts (reference ``training.py:3722-3745``)."""
    args = get_args()
    if not getattr(args, "log_progress", False) or not args.save:
        return
    if barrier and dist.is_initialized():
        dist.barrier()
    if not dist.is_initialized() or dist.get_rank() == 0:
        os.makedirs(args.save, exist_ok=True)
        job = os.environ.get("SLURM_JOB_ID", os.environ.get("TORCHELASTIC_RUN_ID", "-"))
        with open(os.path.join(args.save, "progress.txt"), "a") as f:
            f.write(f"{time.strftime('%Y-%m-%d %H:%M:%S')}\tJob ID: {job}\t# GPUs: {dist.get_world_size() if dist.is_initialized() else 1}\t{string}\n")


def destroy_global_state() -> None:
    """Everything a re-entry of ``pretrain`` in the same process must not inherit (reference ``training.py:694-720``, used by the in-process restart): global
    args / timers / writers, the micro-batch calculator, the rerun state machine, model-parallel groups and the async-checkpoint queue."""
    from ..core import parallel_state as ps
    from ..core.rerun_state_machine import destroy_rerun_state_machine
    from . import checkpointing
    from .global_vars import destroy_global_vars

    try:
        checkpointing.maybe_finalize_async_save(blocking=True)
    except Exception:
        pass
    destroy_global_vars()
    try:
        from ..core.num_microbatches_calculator import destroy_num_microbatches_calculator

        destroy_num_microbatches_calculator()
    except ImportError:
        pass
    destroy_rerun_state_machine()
    ps.destroy_model_parallel()


# ---------------------------------------------------------------------------------------------------------
def train_step(forward_step_func, data_iterator, model, optimizer, opt_param_scheduler, config):
    """One optimizer step, wrapped by the rerun state machine (fault attribution, reference :3010-3260)."""
    args, timers = get_args(), get_timers()
    rerun = get_rerun_state_machine()
    while rerun.should_run_forward_backward(data_iterator):
        for m in model:
            m.zero_grad_buffer()
        optimizer.zero_grad()
        fb = get_forward_backward_func()
        losses_reduced = fb(forward_step_func=forward_step_func, data_iterator=data_iterator, model=model if len(model) > 1 else model[0],
                            num_microbatches=get_num_microbatches(), seq_length=args.seq_length, micro_batch_size=args.micro_batch_size, forward_only=False)
    should_checkpoint, should_exit, exit_code = rerun.should_checkpoint_and_exit()
    if should_exit:
        return {}, True, should_checkpoint, should_exit, exit_code, None, None
    empty_level = getattr(args, "empty_unused_memory_level", 0) or 0
    if empty_level >= 1 and torch.cuda.is_available():
        torch.cuda.empty_cache()                       # after forward / backward (reference :3149)
    timers("optimizer", log_level=1).start(barrier=args.timing_log_level > 1)
    update_successful, grad_norm, num_zeros = optimizer.step()
    timers("optimizer").stop()
    if empty_level >= 2 and torch.cuda.is_available():
        torch.cuda.empty_cache()                       # after the optimizer step too (reference :3229)
    if update_successful:
        opt_param_scheduler.step(increment=get_num_microbatches() * args.micro_batch_size * args.data_parallel_size)
        skipped = 0
    else:
        skipped = 1
    loss_reduced = {}
    if ps.is_pipeline_last_stage(ignore_virtual=True) and losses_reduced:
        for key in losses_reduced[0]:
            vals = torch.stack([d[key].float().reshape(()) for d in losses_reduced])
            v = vals.mean()
            if ps.get_data_parallel_world_size(with_context_parallel=True) > 1:
                dist.all_reduce(v, group=ps.get_data_parallel_group(with_context_parallel=True))
                v = v / ps.get_data_parallel_world_size(with_context_parallel=True)
            loss_reduced[key] = v
    return loss_reduced, skipped, should_checkpoint, should_exit, exit_code, grad_norm, num_zeros


def training_log(loss_dict, total_loss_dict, learning_rate, iteration, loss_scale, grad_norm, elapsed_per_iter, model_flops_per_iter):
    args = get_args()
    for k, v in loss_dict.items():
        total_loss_dict[k] = total_loss_dict.get(k, 0.0) + float(v)
    if iteration % args.log_interval != 0:
        return
    world = dist.get_world_size() if dist.is_initialized() else 1
    tput = model_flops_per_iter / (elapsed_per_iter * 1e12 * world)
    tok_s = get_current_global_batch_size() * args.seq_length / elapsed_per_iter
    s = f" iteration {iteration:8d}/{args.train_iters:8d} | consumed samples: {args.consumed_train_samples:12d} |"
    s += f" elapsed time per iteration (ms): {elapsed_per_iter * 1000.0:.1f} | throughput per GPU (TFLOP/s/GPU): {tput:.1f} | tokens/s: {tok_s:.0f} |"
    s += f" learning rate: {learning_rate:.6E} | global batch size: {get_current_global_batch_size():5d} |"
    logged = {}
    for k in list(total_loss_dict):
        logged[k] = total_loss_dict[k] / args.log_interval
        s += f" {k}: {logged[k]:.6E} |"
        total_loss_dict[k] = 0.0
    s += f" loss scale: {float(loss_scale):.1f} |"
    if grad_norm is not None:
        s += f" grad norm: {float(grad_norm):.3f} |"
    if torch.cuda.is_available():
        s += f" mem-max-allocated-GiB: {torch.cuda.max_memory_allocated() / 2**30:.2f} |"
    print_rank_last(s)
    g = lambda n, d=None: getattr(args, n, d)  # noqa: E731
    tb_every = g("tensorboard_log_interval", 1) or 1
    for w in (_GLOBALS.get("tensorboard"), _GLOBALS.get("wandb")):
        if w is None or iteration % tb_every != 0:
            continue
        scal = {"iteration-time": elapsed_per_iter, "throughput": tput, "tokens-per-sec": tok_s, "learning-rate": learning_rate,
                "batch-size": get_current_global_batch_size(), **{k: float(v) for k, v in logged.items()}}
        if g("log_loss_scale_to_tensorboard", True):
            scal["loss-scale"] = float(loss_scale)
        if g("log_world_size_to_tensorboard", False):
            scal["world-size"] = world
        if g("log_memory_to_tensorboard", False) and torch.cuda.is_available() and iteration % (g("log_memory_interval") or 1) == 0:
            st = torch.cuda.memory_stats()
            scal.update({"mem-reserved-bytes": st["reserved_bytes.all.current"], "mem-allocated-bytes": st["allocated_bytes.all.current"],
                         "mem-max-allocated-bytes": st["allocated_bytes.all.peak"], "mem-allocated-count": st["allocation.all.current"]})
        if g("log_timers_to_tensorboard", False) and hasattr(w, "add_scalar") and get_timers()._timers:
            get_timers().write(list(get_timers()._timers), w, iteration, normalizer=args.log_interval, reset=False)
        if grad_norm is not None:
            scal["grad-norm"] = float(grad_norm)
        if hasattr(w, "add_scalar"):
            for k, v in scal.items():
                w.add_scalar(k, v, iteration)
        else:
            w.log(scal, step=iteration)
    get_timers().log(normalizer=args.log_interval)


def evaluate(forward_step_func, data_iterator, model, eval_iters: int) -> Dict[str, float]:
    args = get_args()
    for m in model:
        m.eval()
    total: Dict[str, float] = {}
    fb = get_forward_backward_func()
    with torch.no_grad():
        for _ in range(eval_iters):
            out = fb(forward_step_func=forward_step_func, data_iterator=data_iterator, model=model if len(model) > 1 else model[0],
                     num_microbatches=get_num_microbatches(), seq_length=args.seq_length, micro_batch_size=args.micro_batch_size, forward_only=True)
            if ps.is_pipeline_last_stage(ignore_virtual=True):
                for d in out:
                    for k, v in d.items():
                        total[k] = total.get(k, 0.0) + float(v) / (eval_iters * len(out))
    for m in model:
        m.train()
    return total


def train(forward_step_func, model, optimizer, opt_param_scheduler, train_data_iterator, valid_data_iterator, config):
    args, timers = get_args(), get_timers()
    for m in model:
        m.train()
    iteration = args.iteration
    args.consumed_train_samples = getattr(args, "consumed_train_samples", iteration * args.global_batch_size)
    config.finalize_model_grads_func = finalize_model_grads
    config.grad_scale_func = optimizer.scale_loss if args.fp16 else None
    config.timers = timers if args.timing_log_level > 0 else None
    config.no_sync_func = model[0].no_sync if len(model) == 1 else [m.no_sync for m in model]
    if args.overlap_grad_reduce and args.pipeline_model_parallel_size > 1:
        config.grad_sync_func = model[0].start_grad_sync if len(model) == 1 else [m.start_grad_sync for m in model]
    total_loss_dict: Dict[str, float] = {}
    exit_flag = {"sig": False}
    exit_sig = getattr(args, "exit_signal", None) or "SIGTERM"                 # --exit-signal NAME (reference TrainingConfig.exit_signal), SIGTERM by default
    exit_sig = getattr(signal, exit_sig if str(exit_sig).startswith("SIG") else f"SIG{exit_sig}") if isinstance(exit_sig, str) else exit_sig
    signal.signal(exit_sig, lambda *_: exit_flag.__setitem__("sig", True))
    skip_iters = set(getattr(args, "iterations_to_skip", None) or [])
    args.skipped_train_samples = getattr(args, "skipped_train_samples", 0)
    flops_per_iter = num_floating_point_operations(
        num_layers=args.num_layers, hidden_size=args.hidden_size, ffn_hidden_size=args.ffn_hidden_size, num_attention_heads=args.num_attention_heads,
        num_query_groups=args.num_query_groups, kv_channels=args.kv_channels or args.hidden_size // args.num_attention_heads,
        vocab_size=getattr(args, "padded_vocab_size", args.vocab_size), seq_length=args.seq_length, batch_size=args.global_batch_size, swiglu=args.swiglu,
        num_moe_experts=args.num_experts, moe_router_topk=args.moe_router_topk,
    )
    append_to_progress_log(f"Starting job\titeration: {iteration}\tFLOPs so far: {args.num_floating_point_operations_so_far:.4e}")
    straggler = StragglerDetector()
    straggler.configure(dist.get_world_size() if dist.is_initialized() else 1, dist.get_rank() if dist.is_initialized() else 0, enabled=args.log_straggler)
    if args.manual_gc:
        gc.disable()
        gc.collect()
    from . import ft_integration

    one_logger = _GLOBALS.get("one_logger")
    if one_logger is not None:
        one_logger.on_train_start(iteration, args.consumed_train_samples, args.train_iters * args.global_batch_size, args.seq_length, args.train_iters, args.save,
                                  args.async_save, True, args.num_floating_point_operations_so_far)
    stat_loggers = []
    if args.log_activations_interval or args.log_dgrad_interval or args.log_wgrad_interval:
        from .activation_logging import ActivationLogger, DgradLogger, WgradLogger

        out_dir = args.activation_log_dir or os.path.join(args.save or ".", "activation_logs")
        rk = dist.get_rank() if dist.is_initialized() else 0
        stat_loggers = [cls(model, out_dir, iv, rank=rk) for cls, iv in ((ActivationLogger, args.log_activations_interval), (DgradLogger, args.log_dgrad_interval),
                                                                       (WgradLogger, args.log_wgrad_interval)) if iv]
    from ..core import telemetry

    if getattr(args, "trace_spans", None):
        telemetry.set_recorder(telemetry.SpanRecorder(path=args.trace_spans))
    metrics = telemetry.TrainingMetrics(prometheus_port=getattr(args, "prometheus_port", None)) if getattr(args, "prometheus_port", None) else None
    torch_prof = None
    if getattr(args, "record_memory_history", False) and torch.cuda.is_available():
        torch.cuda.memory._record_memory_history(max_entries=100000)
    t_start = time.time()
    t_log = time.time()
    prof_ranks = getattr(args, "profile_ranks", None) or []
    profile_here = not prof_ranks or (dist.get_rank() if dist.is_initialized() else 0) in prof_ranks       # --profile-ranks: empty = every rank
    while iteration < args.train_iters:
        if getattr(args, "use_pytorch_profiler", False) and profile_here and iteration == args.profile_step_start and torch_prof is None:
            acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
            torch_prof = torch.profiler.profile(activities=acts, record_shapes=args.pytorch_profiler_collect_shapes or getattr(args, "record_shapes", False),
                                                with_stack=args.pytorch_profiler_collect_callstack)
            torch_prof.__enter__()
        ft_integration.on_training_step_start()
        if _GLOBALS.get("fault_injector") is not None:
            _GLOBALS["fault_injector"].on_iteration(iteration + 1)
        t_iter = time.time()
        for sl in stat_loggers:
            sl.begin_iteration(iteration + 1)
        if args.profile and profile_here and iteration == args.profile_step_start and torch.cuda.is_available():
            torch.cuda.cudart().cudaProfilerStart()
        update_num_microbatches(args.consumed_train_samples, consistency_check=True)
        if (iteration + 1) in skip_iters:
            # fast-forward the data iterator by one global batch and move on: no forward, no update (reference ``--iterations-to-skip`` :4648)
            for _ in range(get_num_microbatches()):
                if train_data_iterator is not None:
                    next(train_data_iterator)
            iteration += 1
            args.consumed_train_samples += get_current_global_batch_size()
            args.skipped_train_samples += get_current_global_batch_size()
            ft_integration.on_training_step_end()
            continue
        with straggler(), telemetry.span("train.iteration", iteration=iteration + 1):
            loss_dict, skipped, should_ckpt, should_exit, exit_code, grad_norm, _ = train_step(forward_step_func, train_data_iterator, model, optimizer, opt_param_scheduler, config)
        if should_ckpt and args.save:
            checkpointing.save_checkpoint(iteration, model, optimizer, opt_param_scheduler, args.save, vars_for_ckpt(args), args.num_floating_point_operations_so_far,
                                          rerun_state=get_rerun_state_machine().state_dict(None, False))
        if should_exit:
            sys.exit(exit_code)
        iteration += 1
        if getattr(args, "train_sync_interval", None) and iteration % args.train_sync_interval == 0 and torch.cuda.is_available():
            torch.cuda.synchronize()                     # keep the host from running ahead of the device (reference :3972)
        ft_integration.on_training_step_end()
        for sl in stat_loggers:
            if hasattr(sl, "collect") and sl.handles:
                sl.collect()
            sl.end_iteration()
        if one_logger is not None:
            one_logger.track_iteration(time.time() - t_iter, get_current_global_batch_size(), args.seq_length, flops_per_iter)
        if metrics is not None:
            metrics.record_iteration(iteration_time_s=time.time() - t_iter, tokens=get_current_global_batch_size() * args.seq_length, flops=flops_per_iter,
                                     world=dist.get_world_size() if dist.is_initialized() else 1, grad_norm=float(grad_norm) if grad_norm is not None else None)
        if torch_prof is not None and iteration == args.profile_step_end:
            torch_prof.__exit__(None, None, None)
            out_dir = os.path.join(os.path.dirname(os.path.abspath(args.tensorboard_dir)) if args.tensorboard_dir else (args.save or "."), "torch_profile")
            os.makedirs(out_dir, exist_ok=True)
            torch_prof.export_chrome_trace(os.path.join(out_dir, f"rank-{dist.get_rank() if dist.is_initialized() else 0}.json.gz"))
            torch_prof = None
        args.consumed_train_samples += get_current_global_batch_size()
        args.num_floating_point_operations_so_far += flops_per_iter
        if iteration % args.log_interval == 0:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            el = (time.time() - t_log) / args.log_interval
            t_log = time.time()
            lr = max((g["lr"] for g in optimizer.param_groups), default=0.0)
            training_log(loss_dict, total_loss_dict, lr, iteration, optimizer.get_loss_scale(), grad_norm, el, flops_per_iter)
            if args.log_straggler:
                straggler.report(flops_per_iter * args.log_interval, args.log_interval)
        else:
            for k, v in loss_dict.items():
                total_loss_dict[k] = total_loss_dict.get(k, 0.0) + float(v)
        if args.profile and profile_here and iteration == args.profile_step_end and torch.cuda.is_available():
            torch.cuda.cudart().cudaProfilerStop()
        if args.eval_interval and args.eval_iters and iteration % args.eval_interval == 0 and valid_data_iterator is not None and \
                (getattr(args, "start_eval_at_iter", None) is None or iteration >= args.start_eval_at_iter):
            gc_eval = args.manual_gc and getattr(args, "manual_gc_eval", True)
            if gc_eval:
                gc.collect()                            # collect before / after evaluation so its garbage does not land inside a training step
            res = evaluate(forward_step_func, valid_data_iterator, model, args.eval_iters)
            if gc_eval:
                gc.collect(generation=0)
            for w in (_GLOBALS.get("tensorboard"), _GLOBALS.get("wandb")):
                if w is None:
                    continue
                vals = {f"{k} validation": v for k, v in res.items()}
                if getattr(args, "log_validation_ppl_to_tensorboard", False):
                    vals.update({f"{k} validation ppl": math.exp(min(20.0, v)) for k, v in res.items()})
                if hasattr(w, "add_scalar"):
                    for k, v in vals.items():
                        w.add_scalar(k, v, iteration)
                else:
                    w.log(vals, step=iteration)
            print_rank_last(f" validation loss at iteration {iteration} | " + " | ".join(f"{k}: {v:.6E}" for k, v in res.items()))
        iv = getattr(args, "check_weight_hash_across_dp_replicas_interval", None)
        if iv and iteration % iv == 0:
            from ..core.utils import check_param_hashes_across_dp_replicas

            assert check_param_hashes_across_dp_replicas(model, cross_check=True), f"parameter hashes differ across data-parallel replicas at iteration {iteration}"
        if args.manual_gc and args.manual_gc_interval and iteration % args.manual_gc_interval == 0:
            gc.collect()
        checkpointing.maybe_finalize_async_save(blocking=False)
        if getattr(args, "non_persistent_save_interval", None) and iteration % args.non_persistent_save_interval == 0:
            kind = getattr(args, "non_persistent_ckpt_type", None) or ("local" if getattr(args, "non_persistent_local_ckpt_dir", None) else None)
            if kind is not None:
                checkpointing.save_non_persistent_checkpoint(iteration, model, optimizer, opt_param_scheduler, kind, args.save, getattr(args, "non_persistent_global_ckpt_dir", None),
                                                             getattr(args, "non_persistent_local_ckpt_dir", None), vars_for_ckpt(args), args.num_floating_point_operations_so_far,
                                                             async_save=args.async_save, local_algo=getattr(args, "non_persistent_local_ckpt_algo", "fully_parallel"),
                                                             replication=True if getattr(args, "replication", False) else None, replication_jump=getattr(args, "replication_jump", None),
                                                             replication_factor=getattr(args, "replication_factor", 2) or 2)
        saved = False
        if args.save and args.save_interval and iteration % args.save_interval == 0:
            checkpointing.save_checkpoint(iteration, model, optimizer if not args.no_save_optim else None, opt_param_scheduler, args.save, vars_for_ckpt(args),
                                          args.num_floating_point_operations_so_far, async_save=args.async_save, keep_last=args.keep_last_checkpoints,
                                          assume_constant_structure=getattr(args, 'ckpt_assume_constant_structure', False), retain_interval=getattr(args, "save_retain_interval", None))
            saved = True
            append_to_progress_log(f"Saved checkpoint\titeration: {iteration}\tFLOPs so far: {args.num_floating_point_operations_so_far:.4e}\ttokens so far: {args.consumed_train_samples * args.seq_length:.4e}",
                                   barrier=False)
        stop = exit_flag["sig"] or (args.exit_interval and iteration % args.exit_interval == 0) or \
            (args.exit_duration_in_mins and (time.time() - t_start) / 60.0 > args.exit_duration_in_mins)
        if dist.is_initialized() and (exit_flag["sig"] or args.exit_duration_in_mins):
            t = torch.tensor([1 if stop else 0], device="cuda" if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            stop = bool(t.item())
        if stop:
            if args.save and not saved:
                checkpointing.save_checkpoint(iteration, model, optimizer, opt_param_scheduler, args.save, vars_for_ckpt(args), args.num_floating_point_operations_so_far)
            print_rank_0(f"exiting program at iteration {iteration}")
            break
    checkpointing.maybe_finalize_async_save(blocking=True)
    if getattr(args, "record_memory_history", False) and torch.cuda.is_available():
        torch.cuda.memory._dump_snapshot(args.memory_snapshot_path)
        torch.cuda.memory._record_memory_history(enabled=None)
    if one_logger is not None:
        one_logger.on_train_end()
    for w in (_GLOBALS.get("tensorboard"), _GLOBALS.get("wandb")):
        if w is not None and hasattr(w, "flush"):
            w.flush()
    args.iteration = iteration
    return iteration


def vars_for_ckpt(args) -> dict:
    return {k: v for k, v in vars(args).items() if isinstance(v, (int, float, str, bool, type(None), list, tuple))}


def get_train_valid_test_num_samples(args) -> List[int]:
    """How many samples each split must be able to serve for this run (reference ``training.get_train_valid_test_num_samples``): the dataset builders size
    their shuffled index caches from these numbers, which is why ``tools/prepare_cache.py`` must compute them exactly like the training job."""
    n_train = args.train_iters * args.global_batch_size
    n_eval = (args.train_iters // max(args.eval_interval, 1) + 1) * args.eval_iters * args.global_batch_size
    return [n_train, n_eval, args.eval_iters * args.global_batch_size]


def pretrain(train_valid_test_dataset_provider: Callable, model_provider: Callable, forward_step_func: Callable, argv=None,
             extra_args_provider=None, args_defaults: Optional[dict] = None):
    """Main entry (reference ``pretrain`` :1500).  ``dataset_provider(num_samples[3]) -> (train, valid, test)``
    datasets; ``model_provider(pre_process, post_process, vp_stage) -> module``;
    ``forward_step_func(data_iterator, model) -> (output, loss_func)``."""
    args = initialize_megatron(argv, extra_args_provider, args_defaults)
    checkpointing.configure(verify_integrity=getattr(args, "verify_integrity", False))
    model, optimizer, scheduler = setup_model_and_optimizer(model_provider)
    config = model[0].module.config if hasattr(model[0], "module") else model[0].config
    train_ds, valid_ds, test_ds = train_valid_test_dataset_provider(get_train_valid_test_num_samples(args))
    from .data import build_pretraining_data_loader

    consumed = getattr(args, "consumed_train_samples", None)
    if consumed is None:
        consumed = args.iteration * args.global_batch_size
    train_it = RerunDataIterator(iter(build_pretraining_data_loader(train_ds, consumed, args))) if train_ds is not None else None
    valid_it = iter(build_pretraining_data_loader(valid_ds, 0, args)) if valid_ds is not None else None
    print_rank_0("training ...")
    if args.train_iters > 0 and not getattr(args, "skip_train", False):
        train(forward_step_func, model, optimizer, scheduler, train_it, valid_it, config)
    if args.eval_iters and valid_it is not None:
        res = evaluate(forward_step_func, valid_it, model, args.eval_iters)
        print_rank_last(" final validation | " + " | ".join(f"{k}: {v:.6E}" for k, v in res.items()))
    if args.save and args.iteration and (not args.save_interval or args.iteration % args.save_interval != 0):
        checkpointing.save_checkpoint(args.iteration, model, optimizer, scheduler, args.save, vars_for_ckpt(args), args.num_floating_point_operations_so_far)
    return model, optimizer
