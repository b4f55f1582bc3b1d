"""Training-level checkpoint save / load (reference ``megatron/training/checkpointing.py:605,2452``).

Layout: ``<save>/iter_{it:07d}/`` = a ``torch_dist`` distributed checkpoint (``dist_checkpointing.save``)
whose state dict has the reference's top-level keys (``args``, ``checkpoint_version``, ``iteration``,
``model`` / ``model{i}``, ``optimizer``, ``opt_param_scheduler``, ``rng_state``, ``rerun_state_machine``,
``num_floating_point_operations_so_far``); ``<save>/latest_checkpointed_iteration.txt`` is the tracker.
"""
from __future__ import annotations

import os
import random
import shutil
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..core import dist_checkpointing, parallel_state as ps
from ..core.dist_checkpointing.mapping import ShardedObject
from ..core.dist_checkpointing.strategies.async_utils import AsyncCallsQueue
from ..core.dist_checkpointing.strategies.fully_parallel import FullyParallelSaveStrategyWrapper
from ..core.tensor_parallel.random import get_cuda_rng_tracker

TRACKER = "latest_checkpointed_iteration.txt"
_ASYNC_QUEUE = AsyncCallsQueue()


def get_checkpoint_name(checkpoints_path: str, iteration: int, release: bool = False) -> str:
    return os.path.join(checkpoints_path, "release" if release else f"iter_{iteration:07d}")


def get_checkpoint_tracker_filename(checkpoints_path: str) -> str:
    return os.path.join(checkpoints_path, TRACKER)


def read_metadata(tracker_filename: str) -> Tuple[int, bool]:
    with open(tracker_filename) as f:
        s = f.read().strip()
    if s == "release":
        return 0, True
    return int(s), False


def _rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def get_rng_state(data_parallel_random_init: bool = False) -> ShardedObject:
    """RNG state of this (pp, tp) position as a ShardedObject (reference :451-511)."""
    state = {
        "random_rng_state": random.getstate(),
        "np_rng_state": np.random.get_state(),
        "torch_rng_state": torch.get_rng_state(),
        "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
        "rng_tracker_states": get_cuda_rng_tracker().get_states(),
    }
    pp_rank, pp = ps.get_pipeline_model_parallel_rank(), ps.get_pipeline_model_parallel_world_size()
    tp_rank, tp = ps.get_tensor_model_parallel_rank(), ps.get_tensor_model_parallel_world_size()
    dp_rank = ps.get_data_parallel_rank(with_context_parallel=True)
    return ShardedObject("rng_state", [state], (pp, tp), (pp_rank, tp_rank), replica_id=dp_rank)


def set_rng_state(states) -> None:
    s = states[0] if isinstance(states, list) else states
    random.setstate(s["random_rng_state"])
    np.random.set_state(s["np_rng_state"])
    torch.set_rng_state(s["torch_rng_state"])
    if s.get("cuda_rng_state") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state(s["cuda_rng_state"])
    if s.get("rng_tracker_states"):
        get_cuda_rng_tracker().set_states(s["rng_tracker_states"])


def generate_state_dict(args: Dict[str, Any], model: List, optimizer, opt_param_scheduler, iteration: int, rng_state=None,
                        optim_sd_kwargs: Optional[dict] = None, num_floating_point_operations_so_far: float = 0.0, rerun_state=None, is_loading=False):
    sd: Dict[str, Any] = {"args": args, "checkpoint_version": 3.0, "iteration": iteration,
                          "num_floating_point_operations_so_far": num_floating_point_operations_so_far}
    for i, m in enumerate(model):
        key = "model" if len(model) == 1 else f"model{i}"
        sd[key] = m.sharded_state_dict(prefix="")
    if optimizer is not None and not getattr(optimizer, "is_stub_optimizer", False):
        model_sd = sd["model"] if len(model) == 1 else {k: v for i in range(len(model)) for k, v in sd[f"model{i}"].items()}
        sd["optimizer"] = optimizer.sharded_state_dict(model_sd, is_loading=is_loading, metadata=optim_sd_kwargs or {"distrib_optim_sharding_type": "fully_reshardable"})
    if opt_param_scheduler is not None:
        sd["opt_param_scheduler"] = opt_param_scheduler.state_dict()
    if rng_state is not None:
        sd["rng_state"] = rng_state
    if rerun_state is not None:
        sd["rerun_state_machine"] = rerun_state
    return sd


def save_checkpoint(iteration: int, model: List, optimizer, opt_param_scheduler, save_dir: str, args: Optional[Dict[str, Any]] = None,
                    num_floating_point_operations_so_far: float = 0.0, async_save: bool = False, fully_parallel_save: bool = True,
                    keep_last: Optional[int] = None, rerun_state=None, optim_sharding_type: str = "fully_reshardable"):
    """Collective over all ranks.  Returns after the checkpoint is durable (or, with ``async_save``,
    after staging; call ``maybe_finalize_async_save`` from the training loop)."""
    ckpt = get_checkpoint_name(save_dir, iteration)
    if _rank() == 0:
        os.makedirs(save_dir, exist_ok=True)
        if os.path.isdir(ckpt):
            shutil.rmtree(ckpt)
    if dist.is_initialized():
        dist.barrier()
    sd = generate_state_dict(dict(args or {}), model, optimizer, opt_param_scheduler, iteration, get_rng_state(),
                             {"distrib_optim_sharding_type": optim_sharding_type}, num_floating_point_operations_so_far, rerun_state)
    strategy = None
    if fully_parallel_save and ps.is_initialized() and ps.get_data_parallel_world_size(with_context_parallel=True) > 1:
        strategy = FullyParallelSaveStrategyWrapper(None, ps.get_data_parallel_group(with_context_parallel=True))

    def write_tracker():
        if _rank() == 0:
            with open(get_checkpoint_tracker_filename(save_dir), "w") as f:
                f.write(str(iteration))
            if keep_last:
                _cleanup_old(save_dir, keep_last)

    req = dist_checkpointing.save(sd, ckpt, sharded_strategy=strategy, async_sharded_save=async_save)
    if req is not None:
        req.add_finalize_fn(write_tracker)
        _ASYNC_QUEUE.schedule_async_request(req)
    else:
        write_tracker()
        if dist.is_initialized():
            dist.barrier()
    return ckpt


def maybe_finalize_async_save(blocking: bool = False):
    return _ASYNC_QUEUE.maybe_finalize_async_calls(blocking=blocking, no_dist=not dist.is_initialized())


def _cleanup_old(save_dir: str, keep_last: int):
    its = sorted(int(d[5:]) for d in os.listdir(save_dir) if d.startswith("iter_") and d[5:].isdigit())
    for it in its[:-keep_last]:
        shutil.rmtree(get_checkpoint_name(save_dir, it), ignore_errors=True)


def load_checkpoint(model: List, optimizer, opt_param_scheduler, load_dir: str, iteration: Optional[int] = None, load_optim: bool = True,
                    load_rng: bool = True, strict: bool = True, optim_sharding_type: Optional[str] = None) -> Tuple[int, float]:
    """Returns ``(iteration, num_floating_point_operations_so_far)``; (0, 0) when nothing to load.
    TP/PP/DP may differ from the run that saved (resharding happens in ``dist_checkpointing.load``)."""
    tracker = get_checkpoint_tracker_filename(load_dir)
    if iteration is None:
        if not os.path.isfile(tracker):
            return 0, 0.0
        iteration, release = read_metadata(tracker)
    ckpt = get_checkpoint_name(load_dir, iteration)
    common = dist_checkpointing.load_common_state_dict(ckpt)
    saved_type = (common.get("optimizer") or {}).get("param_state_sharding_type") if isinstance(common.get("optimizer"), dict) else None
    kind = optim_sharding_type or saved_type or "fully_reshardable"
    sd = generate_state_dict({}, model, optimizer if load_optim else None, opt_param_scheduler, iteration,
                             get_rng_state() if load_rng else None, {"distrib_optim_sharding_type": kind}, is_loading=True)
    sd.pop("args", None)
    if not load_rng:
        sd.pop("rng_state", None)
    elif common_rng_shape_changed(ckpt, sd):
        sd.pop("rng_state", None)  # parallel layout changed: RNG streams cannot be mapped
    loaded = dist_checkpointing.load(sd, ckpt)
    for i, m in enumerate(model):
        key = "model" if len(model) == 1 else f"model{i}"
        m.load_state_dict(loaded[key], strict=strict)
    if optimizer is not None and load_optim and "optimizer" in loaded and not getattr(optimizer, "is_stub_optimizer", False):
        optimizer.load_sharded_state_dict(loaded["optimizer"])
    if opt_param_scheduler is not None and "opt_param_scheduler" in loaded:
        opt_param_scheduler.load_state_dict(loaded["opt_param_scheduler"])
    if load_rng and "rng_state" in loaded and loaded["rng_state"] is not None:
        try:
            set_rng_state(loaded["rng_state"])
        except Exception:
            pass
    return int(loaded.get("iteration", iteration)), float(loaded.get("num_floating_point_operations_so_far", 0.0))


def common_rng_shape_changed(ckpt: str, sd) -> bool:
    try:
        md = dist_checkpointing.strategies.torch_dist.FileSystemReader(ckpt).read_metadata()
    except Exception:
        return False
    want = sd.get("rng_state")
    if want is None:
        return False
    return want.unique_key not in md.state_dict_metadata
