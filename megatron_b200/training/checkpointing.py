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


_VERIFY_INTEGRITY = False


def configure(verify_integrity: bool = False) -> None:
    """``--verify-integrity``: write a SHA-256 manifest (``integrity.json``) with every checkpoint and check it before loading one."""
    global _VERIFY_INTEGRITY
    _VERIFY_INTEGRITY = bool(verify_integrity)


def save_checkpoint(iteration: int, model: List, optimizer, opt_param_scheduler, save_dir: str, args: Optional[Dict[str, Any]] = None,
                    num_floating_point_operations_so_far: float = 0.0, async_save: bool = False, fully_parallel_save: bool = True,
                    keep_last: Optional[int] = None, rerun_state=None, optim_sharding_type: str = "fully_reshardable",
                    assume_constant_structure: bool = False, retain_interval: Optional[int] = None):
    """Collective over all ranks.  Returns after the checkpoint is durable (or, with ``async_save``,
    after staging; call ``maybe_finalize_async_save`` from the training loop).  ``assume_constant_structure`` (``--ckpt-assume-constant-structure``): reuse the
    previous save's plan / metadata when nothing changed structurally (dist_checkpointing SavePlanCache).  ``retain_interval`` (``--save-retain-interval``): once
    the new checkpoint is durable, the PREVIOUS one is deleted unless its iteration is a multiple of the interval (or it is a symbolic link)."""
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
            tracker, prev = get_checkpoint_tracker_filename(save_dir), 0
            if retain_interval and os.path.isfile(tracker):
                text = open(tracker).read().strip()
                prev = int(text) if text.isdigit() else 0
            with open(tracker, "w") as f:
                f.write(str(iteration))
            if keep_last:
                _cleanup_old(save_dir, keep_last)
            if retain_interval and prev > 0 and prev != iteration and prev % retain_interval != 0:
                old = get_checkpoint_name(save_dir, prev)
                if not os.path.islink(old):
                    shutil.rmtree(old, ignore_errors=True)

    req = dist_checkpointing.save(sd, ckpt, sharded_strategy=strategy, async_sharded_save=async_save, cached_structure=assume_constant_structure,
                                   verify_integrity=_VERIFY_INTEGRITY)
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
                    load_rng: bool = True, strict: bool = True, optim_sharding_type: Optional[str] = None, fully_parallel_load: bool = False,
                    dist_ckpt_strictness: Optional[str] = None) -> Tuple[int, float]:
    """Returns ``(iteration, num_floating_point_operations_so_far)``; (0, 0) when nothing to load.
    TP/PP/DP may differ from the run that saved (resharding happens in ``dist_checkpointing.load``).  ``fully_parallel_load`` (``--ckpt-fully-parallel-load``):
    every DP-replicated shard is read from storage by ONE rank of the dp-cp group and exchanged; ``dist_ckpt_strictness``: how key mismatches between the
    checkpoint and the model are treated (``--dist-ckpt-strictness``, the StrictHandling values)."""
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
    kw = {}
    if fully_parallel_load and ps.is_initialized() and ps.get_data_parallel_world_size(with_context_parallel=True) > 1:
        from ..core.dist_checkpointing.strategies.fully_parallel import FullyParallelLoadStrategyWrapper

        kw["sharded_strategy"] = FullyParallelLoadStrategyWrapper(None, ps.get_data_parallel_group(with_context_parallel=True))
    if dist_ckpt_strictness is not None:
        kw["strict"] = dist_ckpt_strictness
    loaded = dist_checkpointing.load(sd, ckpt, verify_integrity=_VERIFY_INTEGRITY, **kw)
    if isinstance(loaded, tuple):            # strictness modes that return the mismatching keys
        loaded, missing, unexpected = loaded
        if _rank() == 0 and (missing or unexpected):
            print(f" > checkpoint key mismatch: {len(missing)} missing, {len(unexpected)} unexpected", flush=True)
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


# ---------------------------------------------------------------------------------------------------------------------------------------------
# checkpoint args → current args   (reference ``load_args_from_checkpoint`` :2199, ``check_checkpoint_args`` :176)
# ---------------------------------------------------------------------------------------------------------------------------------------------
_ARCH_ARGS = ("num_layers", "hidden_size", "ffn_hidden_size", "num_attention_heads", "num_query_groups", "kv_channels", "seq_length", "max_position_embeddings",
              "position_embedding_type", "normalization", "swiglu", "untie_embeddings_and_output_weights", "add_bias_linear", "disable_bias_linear", "vocab_size",
              "padded_vocab_size", "make_vocab_size_divisible_by", "num_experts", "moe_router_topk", "moe_ffn_hidden_size", "multi_latent_attention", "qk_layernorm",
              "rotary_base", "rotary_percent", "tokenizer_type", "group_query_attention")


_MP_ARGS = ("tensor_model_parallel_size", "pipeline_model_parallel_size", "virtual_pipeline_model_parallel_size", "num_layers_per_virtual_pipeline_stage",
            "expert_model_parallel_size", "expert_tensor_parallel_size", "context_parallel_size", "sequence_parallel")


def load_args_from_checkpoint(args, load_dir: Optional[str] = None, iteration: Optional[int] = None, force: bool = True, architecture: bool = True,
                              model_parallel: bool = False):
    """``--use-checkpoint-args``: take the model-architecture arguments from the checkpoint so a run (or an inference server) only needs ``--load``;
    ``--use-mp-args-from-checkpoint-args`` (``model_parallel``): also the model-parallel sizes it was saved with (reference ``checkpointing.py:2355``).
    Returns ``(args, checkpoint_args_dict)``; batch sizes and paths are never overwritten."""
    load_dir = load_dir or getattr(args, "load", None)
    if not load_dir:
        return args, None
    tracker = get_checkpoint_tracker_filename(load_dir)
    if iteration is None:
        if not os.path.isfile(tracker):
            return args, None
        iteration, _ = read_metadata(tracker)
    saved = dist_checkpointing.load_common_state_dict(get_checkpoint_name(load_dir, iteration)).get("args") or {}
    for k in (_ARCH_ARGS if architecture else ()) + (_MP_ARGS if model_parallel else ()):
        if k in saved and saved[k] is not None and (force or getattr(args, k, None) is None):
            setattr(args, k, saved[k])
    return args, saved


def check_checkpoint_args(args, saved: Dict[str, Any], keys=("num_layers", "hidden_size", "num_attention_heads", "num_query_groups", "ffn_hidden_size", "padded_vocab_size",
                                                             "position_embedding_type", "normalization")) -> None:
    """The architecture of a resumed run must match what was saved (parallel sizes may differ: the checkpoint is resharded on load)."""
    bad = {k: (saved[k], getattr(args, k, None)) for k in keys if k in saved and saved[k] is not None and getattr(args, k, None) is not None and saved[k] != getattr(args, k)}
    if bad:
        raise ValueError("checkpoint / command-line architecture mismatch: " + ", ".join(f"{k}: checkpoint {a} vs argument {b}" for k, (a, b) in bad.items()))


# ---------------------------------------------------------------------------------------------------------------------------------------------
# local (non-persistent) checkpoints   (reference ``CheckpointType.LOCAL`` :513, ``checkpointing.py:1834-1905``)
# ---------------------------------------------------------------------------------------------------------------------------------------------
def _local_name(local_dir: str, iteration: int, rank: int) -> str:
    return os.path.join(local_dir, f"iter_{iteration:07d}", f"rank_{rank:05d}.pt")


def save_local_checkpoint(iteration: int, model: List, optimizer, opt_param_scheduler, local_dir: str, keep_last: int = 1, replicate_to_buddy: bool = False,
                          replication_jump: Optional[int] = None, replication_factor: int = 2) -> str:
    """Fast recovery point on NODE-LOCAL storage: every rank dumps its own shards (model, optimizer, scheduler, RNG) with one ``torch.save`` — no global metadata, no
    resharding, so it is as fast as the local disk / ramdisk and only reusable with the same parallel layout.  ``replicate_to_buddy`` also stores the blob of the
    next rank (ring) so a replaced node can be refilled from its neighbour; ``replication_jump`` J / ``replication_factor`` F (``--replication-jump`` /
    ``--replication-factor``): rank n also keeps the blobs of ranks n+J, n+2J, … (F - 1 of them; J = ranks per node puts the replicas on other nodes).  A cluster-wide MIN over the newest complete iteration decides what is loadable."""
    rank = _rank()
    state = {"iteration": iteration, "model": [m.state_dict() for m in model], "rng": get_rng_state().data[0] if hasattr(get_rng_state(), "data") else None,
             "optimizer": optimizer.state_dict() if optimizer is not None and hasattr(optimizer, "state_dict") else None,
             "opt_param_scheduler": opt_param_scheduler.state_dict() if opt_param_scheduler is not None else None,
             "world": dist.get_world_size() if dist.is_initialized() else 1}
    path = _local_name(local_dir, iteration, rank)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)           # atomic: a crash mid-write never leaves a half file under the final name
    if replicate_to_buddy and dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        blob = [None] * world
        dist.all_gather_object(blob, (rank, open(path, "rb").read() if os.path.getsize(path) < (1 << 30) else None))
        jump = replication_jump or 1
        for k in range(1, max(replication_factor, 2)):
            src_rank, data = blob[(rank + k * jump) % world]
            if data is not None and src_rank != rank:
                with open(_local_name(local_dir, iteration, src_rank) + ".buddy", "wb") as f:
                    f.write(data)
    with open(os.path.join(local_dir, f"latest_local_rank_{rank:05d}.txt"), "w") as f:
        f.write(str(iteration))
    its = sorted(int(d.split("_")[1]) for d in os.listdir(local_dir) if d.startswith("iter_"))
    for old in its[:-keep_last] if keep_last else []:
        shutil.rmtree(os.path.join(local_dir, f"iter_{old:07d}"), ignore_errors=True)
    return path


def find_latest_local_checkpoint(local_dir: str) -> int:
    """Newest iteration that EVERY rank can load (−1 when there is none)."""
    rank = _rank()
    mine = -1
    p = os.path.join(local_dir, f"latest_local_rank_{rank:05d}.txt")
    if os.path.isfile(p):
        it = int(open(p).read().strip())
        if os.path.isfile(_local_name(local_dir, it, rank)) or os.path.isfile(_local_name(local_dir, it, rank) + ".buddy"):
            mine = it
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([mine], device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        mine = int(t.item())
    return mine


def load_local_checkpoint(model: List, optimizer, opt_param_scheduler, local_dir: str, iteration: Optional[int] = None) -> int:
    it = find_latest_local_checkpoint(local_dir) if iteration is None else iteration
    if it < 0:
        return -1
    path = _local_name(local_dir, it, _rank())
    if not os.path.isfile(path):
        path += ".buddy"
    state = torch.load(path, map_location="cpu", weights_only=False)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if state.get("world", world) != world:
        raise RuntimeError(f"local checkpoint was written with world size {state['world']}, now {world}: use the global (resharding) checkpoint")
    for m, sd in zip(model, state["model"]):
        m.load_state_dict(sd)
    if optimizer is not None and state.get("optimizer") is not None:
        optimizer.load_state_dict(state["optimizer"])
    if opt_param_scheduler is not None and state.get("opt_param_scheduler") is not None:
        opt_param_scheduler.load_state_dict(state["opt_param_scheduler"])
    if state.get("rng") is not None:
        try:
            set_rng_state(state["rng"])
        except Exception:
            pass
    return it


# ---- non-persistent checkpoints (reference checkpointing.py:513, 1834-1905: CheckpointType.GLOBAL / LOCAL, ``--non-persistent-ckpt-type``) ------------------------
_IN_MEMORY: Dict[str, Any] = {}


def _host_copy(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        return {k: _host_copy(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_host_copy(v) for v in obj)
    return obj


def save_non_persistent_checkpoint(iteration: int, model: List, optimizer, opt_param_scheduler, kind: str, save_dir: Optional[str] = None, global_dir: Optional[str] = None,
                                   local_dir: Optional[str] = None, args: Optional[Dict[str, Any]] = None, num_floating_point_operations_so_far: float = 0.0,
                                   async_save: bool = False, local_algo: str = "fully_parallel", replication: Optional[bool] = None, replication_jump: Optional[int] = None,
                                   replication_factor: int = 2) -> Optional[str]:
    """A frequent recovery point that is NOT part of the persistent series: only the newest one is kept.

    * ``global``: a regular (resharding-capable) distributed checkpoint under ``global_dir`` (default ``<save>/non_persistent``) with its own tracker file;
    * ``local``: every rank dumps its own shards to node-local storage (``save_local_checkpoint``; ``local_algo='fully_parallel'`` additionally replicates each
      blob to the next rank so that a replaced node can be refilled, ``'atomic'`` does not);
    * ``in_memory``: a host copy kept in this process (survives an in-process restart, see ``inprocess_restart.py``)."""
    if kind == "global":
        d = global_dir or os.path.join(save_dir, "non_persistent")
        return save_checkpoint(iteration, model, optimizer, opt_param_scheduler, d, args, num_floating_point_operations_so_far, async_save=async_save, keep_last=1,
                               assume_constant_structure=True)
    if kind == "local":
        assert local_dir, "--non-persistent-local-ckpt-dir is required for local non-persistent checkpoints"
        return save_local_checkpoint(iteration, model, optimizer, opt_param_scheduler, local_dir, replicate_to_buddy=(local_algo == "fully_parallel") if replication is None else replication,
                                     replication_jump=replication_jump, replication_factor=replication_factor)
    if kind == "in_memory":
        _IN_MEMORY.clear()
        _IN_MEMORY.update(iteration=iteration, model=[_host_copy(m.state_dict()) for m in model],
                          optimizer=_host_copy(optimizer.state_dict()) if optimizer is not None and hasattr(optimizer, "state_dict") else None,
                          opt_param_scheduler=opt_param_scheduler.state_dict() if opt_param_scheduler is not None else None,
                          flops=num_floating_point_operations_so_far, rng=get_rng_state().data[0] if hasattr(get_rng_state(), "data") else None)
        return "in_memory"
    raise ValueError(f"unknown non-persistent checkpoint type {kind!r} (global | local | in_memory)")


def load_in_memory_checkpoint(model: List, optimizer, opt_param_scheduler) -> Tuple[int, float]:
    if not _IN_MEMORY:
        return -1, 0.0
    for m, sd in zip(model, _IN_MEMORY["model"]):
        m.load_state_dict(sd)
    if optimizer is not None and _IN_MEMORY.get("optimizer") is not None:
        optimizer.load_state_dict(_IN_MEMORY["optimizer"])
    if opt_param_scheduler is not None and _IN_MEMORY.get("opt_param_scheduler") is not None:
        opt_param_scheduler.load_state_dict(_IN_MEMORY["opt_param_scheduler"])
    if _IN_MEMORY.get("rng") is not None:
        try:
            set_rng_state(_IN_MEMORY["rng"])
        except Exception:
            pass
    return int(_IN_MEMORY["iteration"]), float(_IN_MEMORY["flops"])


def _latest_iteration(ckpt_root: Optional[str]) -> int:
    if not ckpt_root:
        return -1
    t = get_checkpoint_tracker_filename(ckpt_root)
    if not os.path.isfile(t):
        return -1
    it, release = read_metadata(t)
    return -1 if release else it


def load_latest_checkpoint(model: List, optimizer, opt_param_scheduler, load_dir: Optional[str], non_persistent_global_dir: Optional[str] = None,
                           non_persistent_local_dir: Optional[str] = None, pretrained_checkpoint: Optional[str] = None, ckpt_step: Optional[int] = None,
                           exit_on_missing_checkpoint: bool = False, **load_kw) -> Tuple[int, float, str]:
    """The reference's start-up policy (``_load_base_checkpoint``, checkpointing.py:1230-1330): among the persistent series in ``load_dir``, the non-persistent
    global checkpoint and the node-local one, resume from the NEWEST iteration (all ranks agree on it); with nothing to resume from, fall back to
    ``pretrained_checkpoint`` (weights only, iteration 0).  → ``(iteration, flops, source)`` with source in persistent | non_persistent_global | local |
    in_memory | pretrained | none."""
    cands = []
    if ckpt_step is not None:
        cands.append((ckpt_step, "persistent"))
    else:
        cands.append((_latest_iteration(load_dir), "persistent"))
        ng = non_persistent_global_dir or (os.path.join(load_dir, "non_persistent") if load_dir else None)
        cands.append((_latest_iteration(ng), "non_persistent_global"))
        if non_persistent_local_dir:
            cands.append((find_latest_local_checkpoint(non_persistent_local_dir), "local"))
        if _IN_MEMORY:
            cands.append((int(_IN_MEMORY["iteration"]), "in_memory"))
    it, src = max(cands, key=lambda c: c[0])
    if dist.is_initialized() and dist.get_world_size() > 1:
        # filesystem views can differ between nodes: agree on rank 0's choice
        box = [(it, src)]
        dist.broadcast_object_list(box, src=0)
        it, src = box[0]
    if it < 0:
        if pretrained_checkpoint:
            kw = dict(load_kw)
            kw.update(load_optim=False, load_rng=False)
            _, fl = load_checkpoint(model, None, None, pretrained_checkpoint, **kw)
            return 0, 0.0, "pretrained"
        if exit_on_missing_checkpoint:
            if _rank() == 0:
                print(">> '--exit-on-missing-checkpoint' set, and no checkpoint found: exiting", flush=True)
            if dist.is_initialized():
                dist.barrier()
            raise SystemExit(0)
        return 0, 0.0, "none"
    if src == "persistent":
        i, fl = load_checkpoint(model, optimizer, opt_param_scheduler, load_dir, iteration=it if ckpt_step is not None else None, **load_kw)
    elif src == "non_persistent_global":
        i, fl = load_checkpoint(model, optimizer, opt_param_scheduler, non_persistent_global_dir or os.path.join(load_dir, "non_persistent"), **load_kw)
    elif src == "local":
        i, fl = load_local_checkpoint(model, optimizer, opt_param_scheduler, non_persistent_local_dir, iteration=it), 0.0
    else:
        i, fl = load_in_memory_checkpoint(model, optimizer, opt_param_scheduler)
    return i, fl, src
