"""Model-space checkpoints for fully sharded models (reference ``transformer/fsdp_dtensor_checkpoint.py`` + the ``fsdp_dtensor`` checkpoint format).

FSDP keeps, per unit, one flat buffer of all the unit's parameters and gives every rank an equal slice of it — a slice that starts and
ends in the middle of parameters.  Saving those slices verbatim (``fsdp.unit3.master``) ties the checkpoint to the unit boundaries and
the world size.  The model-space format instead names every *parameter*: each rank writes, for every parameter its slice touches, the
piece it owns as a 1-D range ``[start, end)`` of that parameter's flattened values (global shape ``(numel,)``).  The reader asks for any
other set of ranges, so the same files load into

* an FSDP model with a different number of ranks or different unit boundaries,
* a plain (unsharded / DDP) module — ask for whole parameters and reshape,

and optimizer moments, which share the master's layout, ride along under ``<name>.<state>``.  Expert parameters are renamed from the
rank-local expert index to the global one so the expert-parallel size can change too."""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from ..dist_checkpointing.mapping import ShardedTensor

_WRAPPERS = ("module.", "_orig_mod.")


def _strip_wrapper_prefixes(path: str) -> str:
    changed = True
    while changed:
        changed = False
        for w in _WRAPPERS:
            if path.startswith(w):
                path, changed = path[len(w):], True
    return path.replace(".module.", ".")


def _intersect_slice(a: Tuple[int, int], b: Tuple[int, int]) -> Optional[Tuple[int, int]]:
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else None


# ---- expert index bookkeeping ------------------------------------------------------------------------------------------
def get_expert_index_from_key(key: str) -> Optional[int]:
    m = re.search(r"\.local_experts\.(\d+)\.", key) or re.search(r"\.experts\.(\d+)\.", key)
    return int(m.group(1)) if m else None


def expert_param_global_key(key: str, ep_rank: int, num_local_experts: int) -> str:
    """``…experts.local_experts.1.linear_fc1.weight`` on EP rank 2 with 4 local experts → ``…experts.9.linear_fc1.weight``."""
    m = re.search(r"\.local_experts\.(\d+)\.", key)
    if not m:
        return key
    g = ep_rank * num_local_experts + int(m.group(1))
    return key[: m.start()] + f".{g}." + key[m.end():]


def expert_param_local_key(key: str, ep_rank: int, num_local_experts: int) -> Optional[str]:
    """Inverse; ``None`` when the expert lives on another EP rank."""
    m = re.search(r"\.experts\.(\d+)\.", key)
    if not m:
        return key
    g = int(m.group(1))
    if not (ep_rank * num_local_experts <= g < (ep_rank + 1) * num_local_experts):
        return None
    return key[: m.start()] + f".experts.local_experts.{g - ep_rank * num_local_experts}." + key[m.end():]


def handle_experts_in_state_dict(state_dict: Dict[str, object], ep_rank: int, num_local_experts: int) -> Dict[str, object]:
    return {expert_param_global_key(k, ep_rank, num_local_experts): v for k, v in state_dict.items()}


def flatten_state_dict(obj, parent_key: str = "", sep: str = ".") -> Dict[str, object]:
    out = {}
    if isinstance(obj, dict):
        for k, v in obj.items():
            out.update(flatten_state_dict(v, f"{parent_key}{sep}{k}" if parent_key else str(k), sep))
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            out.update(flatten_state_dict(v, f"{parent_key}{sep}{i}" if parent_key else str(i), sep))
    else:
        out[parent_key] = obj
    return out


def print_diff_in_state_dicts(checkpoint_keys: Iterable[str], requested_keys: Iterable[str], limit: int = 100) -> Tuple[List[str], List[str]]:
    ck, rq = set(checkpoint_keys), set(requested_keys)
    missing, unexpected = sorted(rq - ck), sorted(ck - rq)
    for k in missing[:limit]:
        print(f"missing in checkpoint: {k}")
    for k in unexpected[:limit]:
        print(f"not requested:         {k}")
    return missing, unexpected


def get_global_unique_param_name(model_chunks, param) -> Optional[str]:
    chunks = model_chunks if isinstance(model_chunks, (list, tuple)) else [model_chunks]
    for i, m in enumerate(chunks):
        for n, p in m.named_parameters():
            if p is param:
                n = _strip_wrapper_prefixes(n)
                return n if len(chunks) == 1 else f"chunk{i}.{n}"
    return None


# ---- FSDP ⇄ model space ----------------------------------------------------------------------------------------------------
def _param_names(fsdp) -> Dict[int, str]:
    return {id(p): _strip_wrapper_prefixes(n) for n, p in fsdp.module.named_parameters()}


def fsdp_model_space_sharded_state_dict(fsdp, prefix: str = "", extra_states: Optional[Dict[str, List[torch.Tensor]]] = None, ep_rank: int = 0,
                                        num_local_experts: int = 0) -> Dict[str, ShardedTensor]:
    """``{parameter name: this rank's 1-D piece}`` from the fp32 masters (plus ``extra_states[state][unit index]`` tensors laid out like
    the master, e.g. Adam moments → ``<name>.exp_avg``).  A parameter cut by a shard boundary appears on both ranks with different ranges."""
    names = _param_names(fsdp)
    out: Dict[str, ShardedTensor] = {}
    for ui, u in enumerate(fsdp.units):
        lo, hi = u.rank * u.shard_size, (u.rank + 1) * u.shard_size
        off = 0
        for p, n in zip(u.params, u.numels):
            cut = _intersect_slice((off, off + n), (lo, hi))
            if cut is not None:
                name = prefix + names[id(p)]
                if num_local_experts:
                    name = expert_param_global_key(name, ep_rank, num_local_experts)
                sources = {"": u.master.data}
                for sname, per_unit in (extra_states or {}).items():
                    sources["." + sname] = per_unit[ui]
                for suffix, flat in sources.items():
                    piece = flat[cut[0] - lo : cut[1] - lo]
                    key = name + suffix
                    out[key if key not in out else f"{key}#{ui}"] = ShardedTensor(
                        key=key, data=piece, dtype=piece.dtype, local_shape=(cut[1] - cut[0],), global_shape=(n,), global_offset=(cut[0] - off,),
                        axis_fragmentations=None, replica_id=0, allow_shape_mismatch=True)
            off += n
    return out


def load_fsdp_model_space(fsdp, checkpoint_dir: str, prefix: str = "", extra_states: Optional[Dict[str, List[torch.Tensor]]] = None, **kw) -> None:
    """Fill the masters (and ``extra_states``) from a model-space checkpoint written under ANY sharding, then refresh the bf16 shards."""
    from ..dist_checkpointing import serialization

    serialization.load(fsdp_model_space_sharded_state_dict(fsdp, prefix, extra_states, **kw), checkpoint_dir)      # pieces are views: loaded in place
    fsdp.post_optimizer_step()


def load_plain_module_from_model_space(module: torch.nn.Module, checkpoint_dir: str, prefix: str = "", strict: bool = True) -> List[str]:
    """Load the same checkpoint into an unsharded module: whole parameters are requested flat and reshaped."""
    from ..dist_checkpointing import serialization

    want, bufs = {}, {}
    for n, p in module.named_parameters():
        key = prefix + _strip_wrapper_prefixes(n)
        bufs[key] = (p, torch.empty(p.numel(), dtype=torch.float32))
        want[key] = ShardedTensor(key=key, data=bufs[key][1], dtype=torch.float32, local_shape=(p.numel(),), global_shape=(p.numel(),), global_offset=(0,),
                                  axis_fragmentations=None, replica_id=0, allow_shape_mismatch=True)
    meta = serialization.load_tensors_metadata(checkpoint_dir)
    missing = [k for k in want if k not in meta]
    if missing and strict:
        raise KeyError(f"parameters missing from {checkpoint_dir}: {missing[:5]}")
    for k in missing:
        want.pop(k)
    serialization.load(want, checkpoint_dir)
    with torch.no_grad():
        for k in want:
            p, flat = bufs[k]
            p.copy_(flat.view(p.shape).to(p.dtype))
    return missing


def validate_loaded_state_dict(fsdp, reference_full_state: Dict[str, torch.Tensor], atol: float = 0.0) -> List[str]:
    """Names whose gathered value differs from ``reference_full_state`` — a cheap end-to-end check after a resharded load."""
    got = fsdp.gather_full_state_dict()
    return [k for k, v in reference_full_state.items() if torch.is_tensor(v) and k in got and (got[k].float() - v.float()).abs().max().item() > atol]
