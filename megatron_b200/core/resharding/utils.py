"""Data model of the weight-refit path (reference ``megatron/core/resharding/utils.py:19-554``).

Refit = move the weights of a *training* model (one TP/PP/EP layout) into an *inference* model (another layout), every RL
iteration, without going through a checkpoint.  The reference plans this with an LCM micro-tiler per tensor-parallel dimension.
This implementation describes every local shard as **runs**: per axis a list of ``(local_start, global_start, length)`` triples
mapping pieces of the local tensor into the coordinates of the canonical unsharded tensor:

* plain TP                → one run on ``partition_dim``;
* strided TP (``partition_stride`` s, e.g. a gated-linear-unit ``fc1`` = [gate; up]) → s runs;
* block-interleaved TP (``partition_sizes``, e.g. Mamba ``in_proj`` = z, x, B, C, dt) → one run per packed block;
* fused grouped experts   → a run on axis 0 (global expert index = ep_rank * num_local + local).

Planning is then nothing but intersecting runs (``planner.py``), which also covers uneven splits the LCM tiler rejects.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Run = Tuple[int, int, int]                      # (local_start, global_start, length)


@dataclass
class TransferOp:
    """One send or receive of this rank (reference ``utils.py:19``).  ``my_slice`` indexes the local tensor."""
    param_name: str
    peer_rank: int
    is_send: bool
    my_slice: Tuple[slice, ...]
    peer_slice: Tuple[slice, ...]
    task_id: Optional[int] = None
    nbytes: int = 0
    wire_dtype: Optional[torch.dtype] = None     # dtype on the wire = the sender's dtype (the receiver converts)


@dataclass
class ParameterMetadata:
    """What the planner needs to know about one parameter / persistent buffer on one rank (reference ``utils.py:37``)."""
    name: str
    shape: Tuple[int, ...]
    dtype: torch.dtype
    element_size: int
    is_tp: bool = False
    partition_dim: int = 0
    partition_stride: int = 1
    partition_sizes: Optional[List[int]] = None
    is_ep: bool = False
    num_experts: Optional[int] = None
    owner_rank: int = -1
    tensor_parallel_group_ranks: Optional[List[int]] = None
    expert_parallel_group_ranks: Optional[List[int]] = None
    data_parallel_group_ranks: Optional[List[int]] = None
    pipeline_parallel_group_ranks: Optional[List[int]] = None
    resolved_name: Optional[str] = None
    global_expert_index: Optional[int] = None

    # ---- derived -------------------------------------------------------------------------------------------------
    def tp_world(self) -> int:
        return len(self.tensor_parallel_group_ranks) if (self.is_tp and self.tensor_parallel_group_ranks) else 1

    def tp_rank(self) -> int:
        return self.tensor_parallel_group_ranks.index(self.owner_rank) if self.tp_world() > 1 else 0

    def ep_world(self) -> int:
        return len(self.expert_parallel_group_ranks) if (self.is_ep and self.expert_parallel_group_ranks) else 1

    def ep_rank(self) -> int:
        return self.expert_parallel_group_ranks.index(self.owner_rank) if self.ep_world() > 1 else 0

    def is_fused_experts(self) -> bool:
        """Grouped expert tensor [num_local_experts, ...] (axis 0 is the expert axis)."""
        return self.is_ep and self.global_expert_index is None and self.num_experts is not None and len(self.shape) == 3

    def global_shape(self) -> Tuple[int, ...]:
        g = list(self.shape)
        if self.is_tp and self.tp_world() > 1:
            g[self.partition_dim] *= self.tp_world()
        if self.is_fused_experts():
            g[0] *= self.ep_world()
        return tuple(g)

    def runs(self) -> Tuple[Tuple[Run, ...], ...]:
        """Per axis, the runs of this shard in canonical (unsharded) coordinates."""
        out: List[Tuple[Run, ...]] = [((0, 0, n),) for n in self.shape]
        w, r = self.tp_world(), (self.tp_rank() if self.tp_world() > 1 else 0)
        if self.is_tp and w > 1:
            d, local = self.partition_dim, self.shape[self.partition_dim]
            if self.partition_sizes:
                assert sum(self.partition_sizes) == local, f"{self.name}: partition_sizes {self.partition_sizes} != local dim {local}"
                blocks = list(self.partition_sizes)
            else:
                s = max(1, int(self.partition_stride))
                assert local % s == 0, f"{self.name}: local dim {local} not divisible by partition_stride {s}"
                blocks = [local // s] * s
            runs, lo, go = [], 0, 0
            for b in blocks:
                runs.append((lo, go + r * b, b))
                lo += b
                go += b * w
            out[d] = tuple(runs)
        if self.is_fused_experts() and self.ep_world() > 1:
            L = self.shape[0]
            out[0] = ((0, self.ep_rank() * L, L),)
        return tuple(out)


@dataclass
class ShardingDescriptor:
    """One sharded dimension of a parameter between two layouts (reference ``utils.py:88``) — informational here: the run
    intersection does not need it, but the plan report and the tests do."""
    name: str
    dim: int
    src_stride: int
    dst_stride: int
    src_dim_ranks: List[int]
    dst_dim_ranks: List[int]


@dataclass
class ReshardPlan:
    send_ops: List[TransferOp]
    recv_ops: List[TransferOp]
    transform: Optional["object"] = None
    buffer_dtypes: Optional[Dict[str, torch.dtype]] = None
    # bytes this rank sends to / receives from every peer (filled by the planner; sizes the packed per-peer messages)
    send_bytes: Dict[int, int] = field(default_factory=dict)
    recv_bytes: Dict[int, int] = field(default_factory=dict)

    def __str__(self):
        return f"ReshardPlan(sends={len(self.send_ops)}, recvs={len(self.recv_ops)})"


# ---- names ----------------------------------------------------------------------------------------------------------
_EXPERT_RE = re.compile(r"(local_experts\.)(\d+)(\.)|(\.weight|\.bias)(\d+)$")
_LAYER_RE = re.compile(r"((?:^|\.)layers\.)(\d+)(\.)")


def _get_rank_in_group(global_rank: int, group_ranks: Sequence[int]) -> int:
    try:
        return list(group_ranks).index(global_rank)
    except ValueError as e:
        raise ValueError(f"rank {global_rank} is not in group {list(group_ranks)}") from e


def _detect_expert_index_from_param_name(param_name: str) -> Optional[int]:
    m = _EXPERT_RE.search(param_name)
    if m is None:
        return None
    return int(m.group(2) if m.group(2) is not None else m.group(5))


def assign_ep_resolved_name_inplace(meta: ParameterMetadata, *, base_name: Optional[str] = None, fused: bool = False) -> None:
    """Local expert index -> global expert index in ``resolved_name`` (reference ``utils.py:142``): rank 1 of EP=2 with four
    experts names its first expert ``local_experts.0`` although it is global expert 2."""
    name = base_name if base_name is not None else (meta.resolved_name or meta.name)
    meta.resolved_name = name
    if not meta.is_ep:
        return
    local = None if fused else _detect_expert_index_from_param_name(name)
    if local is None or not meta.num_experts:
        return                                           # fused grouped tensor: the expert axis is planned as a run instead
    ep = meta.ep_world()
    per_rank = meta.num_experts // ep
    g = meta.ep_rank() * per_rank + local
    meta.global_expert_index = g

    def sub(m):
        if m.group(2) is not None:
            return f"{m.group(1)}{g}{m.group(3)}"
        return f"{m.group(4)}{g}"
    meta.resolved_name = _EXPERT_RE.sub(sub, name, count=1)


def _build_layer_module_prefix_map(module: torch.nn.Module) -> Dict[str, str]:
    """``decoder.layers.0.`` -> ``decoder.layers.16.`` for every layer module that knows its global ``layer_number``."""
    out: Dict[str, str] = {}
    for mod_name, mod in module.named_modules():
        ln = getattr(mod, "layer_number", None)
        m = re.search(r"(^|\.)layers\.(\d+)$", mod_name)
        if m is None or not isinstance(ln, int):
            continue
        out[mod_name + "."] = mod_name[: m.start(2)] + str(ln - 1) + "."
    return out


def _resolve_global_layer_number_in_name(name: str, prefix_map: Dict[str, str]) -> str:
    best = ""
    for p in prefix_map:
        if name.startswith(p) and len(p) > len(best):
            best = p
    return prefix_map[best] + name[len(best):] if best else name


def assign_resolved_name_inplace(meta: ParameterMetadata, *, layer_module_prefix_map: Optional[Dict[str, str]] = None, base_name: Optional[str] = None, fused: bool = False) -> None:
    """Canonical cross-layout name: global layer numbers (PP) then global expert indices (EP) (reference ``utils.py:183``)."""
    name = base_name if base_name is not None else meta.name
    if layer_module_prefix_map:
        name = _resolve_global_layer_number_in_name(name, layer_module_prefix_map)
    assign_ep_resolved_name_inplace(meta, base_name=name, fused=fused)


# ---- tensors that take part in a refit ---------------------------------------------------------------------------------
def named_persistent_buffers(module: torch.nn.Module) -> Iterator[Tuple[str, torch.Tensor]]:
    """Buffers that are part of the state dict (e.g. the router's expert bias); non-persistent ones are recomputed."""
    for mod_name, mod in module.named_modules():
        skip = getattr(mod, "_non_persistent_buffers_set", set())
        for bname, buf in mod._buffers.items():
            if buf is None or bname in skip:
                continue
            yield (f"{mod_name}.{bname}" if mod_name else bname), buf


def named_refit_tensors(module: torch.nn.Module) -> Iterator[Tuple[str, torch.Tensor]]:
    seen = set()
    for n, p in module.named_parameters():
        if id(p) in seen:
            continue                                      # tied weights travel once
        seen.add(id(p))
        yield n, p
    for n, b in named_persistent_buffers(module):
        if "_extra_state" in n or id(b) in seen:
            continue
        seen.add(id(b))
        yield n, b


_REFIT_CACHE_ATTR = "_mb200_refit_tensor_dict"


def get_refit_tensor_dict(module: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """Name -> tensor, cached on the module: the walk over ``named_modules`` costs more than the copy for small models."""
    d = module.__dict__.get(_REFIT_CACHE_ATTR)
    if d is None:
        d = dict(named_refit_tensors(module))
        module.__dict__[_REFIT_CACHE_ATTR] = d
    return d


def invalidate_refit_tensor_cache(module: torch.nn.Module) -> None:
    module.__dict__.pop(_REFIT_CACHE_ATTR, None)


# ---- metadata extraction ------------------------------------------------------------------------------------------------
def _group_ranks(group, offset: int = 0) -> Optional[List[int]]:
    if group is None:
        return None
    if isinstance(group, (list, tuple)):
        return [int(r) + offset for r in group]
    return [int(r) + offset for r in dist.get_process_group_ranks(group)]


def _gated_fc1_names(module: torch.nn.Module) -> set:
    """Parameters whose local shard is [gate_r; up_r]: planned with stride 2 although the attribute on the tensor says 1
    (this framework keeps the gate/up interleave in the sharded_state_dict factory, like the reference's swiglu factory)."""
    out = set()
    for mod_name, mod in module.named_modules():
        cfg = getattr(mod, "config", None)
        if cfg is None or not getattr(cfg, "gated_linear_unit", False):
            continue
        fc1 = getattr(mod, "linear_fc1", None)
        if fc1 is not None:
            for pn, _ in fc1.named_parameters(recurse=False):
                out.add(f"{mod_name}.linear_fc1.{pn}" if mod_name else f"linear_fc1.{pn}")
        if hasattr(mod, "weight1") and hasattr(mod, "weight2") and hasattr(mod, "num_local_experts"):
            out.add(f"{mod_name}.weight1" if mod_name else "weight1")
    return out


def extract_param_metadata(
    param: torch.Tensor, param_name: str, owner_rank: int, pg_collection=None, num_experts: Optional[int] = None,
    layer_module_prefix_map: Optional[Dict[str, str]] = None, rank_offset: int = 0, *, gated: bool = False,
) -> ParameterMetadata:
    """Reference ``utils.py:309``.  ``pg_collection`` needs ``tp`` / ``ep`` / ``dp`` / ``pp`` (+ ``expt_tp`` / ``expt_dp`` for expert
    parameters); any of them may be a plain list of ranks, which is how the tests and non-collocated callers describe a
    layout that has no process groups in this world.  ``rank_offset`` shifts group-local ranks into the joint world."""
    g = lambda n: getattr(pg_collection, n, None) if pg_collection is not None else None   # noqa: E731
    is_expert = (not getattr(param, "allreduce", True)) or ".experts." in f".{param_name}" or "local_experts." in param_name
    is_tp = bool(getattr(param, "tensor_model_parallel", False))
    dim = int(getattr(param, "partition_dim", 0)) if is_tp else 0
    stride = int(getattr(param, "partition_stride", 1)) if is_tp else 1
    sizes = getattr(param, "partition_sizes", None)
    if gated and is_tp and stride == 1 and sizes is None:
        stride = 2
    tp = g("expt_tp") if (is_expert and g("expt_tp") is not None) else g("tp")
    dp = g("expt_dp") if (is_expert and g("expt_dp") is not None) else g("dp")
    meta = ParameterMetadata(
        name=param_name, shape=tuple(param.shape), dtype=param.dtype, element_size=param.element_size(), is_tp=is_tp, partition_dim=dim,
        partition_stride=stride, partition_sizes=list(sizes) if sizes is not None else None, is_ep=bool(is_expert and num_experts),
        num_experts=num_experts if is_expert else None, owner_rank=owner_rank + rank_offset,
        tensor_parallel_group_ranks=_group_ranks(tp, rank_offset), expert_parallel_group_ranks=_group_ranks(g("ep"), rank_offset) if is_expert else None,
        data_parallel_group_ranks=_group_ranks(dp, rank_offset), pipeline_parallel_group_ranks=_group_ranks(g("pp"), rank_offset),
    )
    if meta.is_tp and meta.tp_world() == 1:
        meta.is_tp = False                                # a TP=1 "shard" is the whole tensor
    # [num_local_experts, ...] grouped tensors (GroupedMLP.weight1 / weight2) carry the expert axis themselves
    assign_resolved_name_inplace(meta, layer_module_prefix_map=layer_module_prefix_map, base_name=param_name, fused=bool(is_expert and param.dim() == 3))
    return meta


def extract_module_metadata(module: Optional[torch.nn.Module], owner_rank: int, pg_collection=None, num_experts: Optional[int] = None, rank_offset: int = 0) -> List[ParameterMetadata]:
    if module is None:
        return []
    prefix_map = _build_layer_module_prefix_map(module)
    gated = _gated_fc1_names(module)
    return [extract_param_metadata(t, n, owner_rank, pg_collection, num_experts, prefix_map, rank_offset, gated=n in gated)
            for n, t in get_refit_tensor_dict(module).items()]


# ---- source selection ----------------------------------------------------------------------------------------------------
def _round_robin_dp(src_meta_list: Sequence[ParameterMetadata], dst_rank: int) -> ParameterMetadata:
    ordered = sorted(src_meta_list, key=lambda m: m.owner_rank)
    return ordered[dst_rank % len(ordered)]


def select_src_metadata_balanced(src_meta_list: Sequence[ParameterMetadata], dst_metadata: ParameterMetadata, dst_rank: int) -> ParameterMetadata:
    """Among data-parallel replicas holding the same shard, take the one on the destination rank itself (no traffic), else
    spread destinations round-robin over the replicas so no single source's NVLink egress is the bottleneck
    (reference ``utils.py:531``)."""
    if not src_meta_list:
        raise ValueError(f"no source metadata for {dst_metadata.resolved_name or dst_metadata.name}")
    for m in src_meta_list:
        if m.owner_rank == dst_rank:
            return m
    return _round_robin_dp(src_meta_list, dst_rank)
