"""Parameter → flat-buffer layout rules shared by DDP and the distributed optimizers (reference ``optimizer/param_layout.py:19-106``).

A layout is computed ONCE (by whoever owns the sharding — the ZeRO-1 optimizer or the layer-wise optimizer) and handed to the DDP
buffer, so "which elements of the buffer belong to shard r" has a single source of truth:

* every parameter starts on a 64-element boundary (128 B for bf16: vectorised kernels and TMA never straddle two parameters),
* every bucket ends on a multiple of ``lcm(dp, 128)`` so a reduce-scatter hands each rank an aligned, equal shard (optionally also a
  multiple of 2^16 elements, which keeps NCCL's ring chunks full),
* buffers are keyed by (param dtype, grad dtype, expert-parallel?, layer-wise-managed?): each key is a physically separate buffer with
  its own index space.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

import torch


def pad_to_divisor(value: int, divisor: int) -> int:
    return -(-value // divisor) * divisor


def pad_param_start(param_start_index: int) -> int:
    return pad_to_divisor(param_start_index, 64)


def bucket_end_divisor(data_parallel_world_size: int, pad_for_high_nccl_busbw: bool) -> int:
    return math.lcm(data_parallel_world_size, 128, 2 ** 16) if pad_for_high_nccl_busbw else math.lcm(data_parallel_world_size, 128)


def pad_bucket_end(bucket_end_index: int, data_parallel_world_size: int, pad_for_high_nccl_busbw: bool) -> int:
    return pad_to_divisor(bucket_end_index, bucket_end_divisor(data_parallel_world_size, pad_for_high_nccl_busbw))


@dataclass(frozen=True)
class BufferKey:
    param_dtype: torch.dtype                  # storage dtype (uint8 for fp8 / nvfp4 parameters)
    grad_dtype: torch.dtype
    is_expert_parallel: bool
    is_managed_by_layer_wise_optimizer: bool = False


@dataclass
class PerBufferParamLayout:
    param_index_map: Dict[torch.nn.Parameter, Tuple[int, int, int]] = field(default_factory=dict)   # param → (start, end, bucket id)
    bucket_indices: List[Tuple[int, int]] = field(default_factory=list)
    per_bucket_numel_unpadded: List[int] = field(default_factory=list)
    param_indices: List[int] = field(default_factory=list)
    num_optimizer_shards: Optional[int] = None

    @property
    def numel(self) -> int:
        return self.bucket_indices[-1][1] if self.bucket_indices else 0

    def shard_range(self, bucket_id: int, shard: int) -> Tuple[int, int]:
        s, e = self.bucket_indices[bucket_id]
        n = (e - s) // self.num_optimizer_shards
        return s + shard * n, s + (shard + 1) * n


@dataclass
class FullParamLayout:
    layouts: Dict[BufferKey, PerBufferParamLayout] = field(default_factory=dict)


def compute_per_buffer_layout(params: Iterable[torch.nn.Parameter], data_parallel_world_size: int, bucket_size: Optional[int] = None,
                              sharded: bool = True, pad_for_high_nccl_busbw: bool = False, whole_params_per_shard: bool = False) -> PerBufferParamLayout:
    """Walk ``params`` in REVERSE (the order backward produces gradients), cutting a bucket whenever ``bucket_size`` elements are reached.

    ``whole_params_per_shard`` (layer-wise optimizer): a parameter never straddles two shards of its bucket — each parameter is placed
    at the start of the next shard that still has room, so every rank can run a whole-matrix update rule on what it owns."""
    out = PerBufferParamLayout(num_optimizer_shards=data_parallel_world_size if sharded else None)
    params = list(params)[::-1]
    offset = bucket_start = 0
    bucket_id = 0
    in_bucket = 0

    def close(end_unpadded: int) -> int:
        nonlocal bucket_start, bucket_id, in_bucket
        out.per_bucket_numel_unpadded.append(end_unpadded - bucket_start)
        end = pad_bucket_end(end_unpadded, data_parallel_world_size, pad_for_high_nccl_busbw) if sharded else pad_param_start(end_unpadded)
        out.bucket_indices.append((bucket_start, end))
        bucket_start, bucket_id, in_bucket = end, bucket_id + 1, 0
        return end

    if whole_params_per_shard and sharded:
        # one bucket; shard capacity = the smallest aligned size that fits a greedy first-fit assignment of whole parameters
        sizes = [pad_param_start(p.numel()) for p in params]
        cap = pad_to_divisor(max(max(sizes, default=0), -(-sum(sizes) // data_parallel_world_size)), 128)
        while True:
            fill, placed, ok = [0] * data_parallel_world_size, [], True
            for sz in sizes:
                r = next((i for i in range(data_parallel_world_size) if fill[i] + sz <= cap), None)
                if r is None:
                    ok = False
                    break
                placed.append((r, fill[r]))
                fill[r] += sz
            if ok:
                break
            cap += 128
        for i, (p, (r, off)) in enumerate(zip(params, placed)):
            out.param_index_map[p] = (r * cap + off, r * cap + off + p.numel(), 0)
            out.param_indices.append(i)
        out.per_bucket_numel_unpadded.append(sum(p.numel() for p in params))
        out.bucket_indices.append((0, cap * data_parallel_world_size))
        return out

    for i, p in enumerate(params):
        start = pad_param_start(offset) if sharded else offset
        end = start + p.numel()
        out.param_index_map[p] = (start, end, bucket_id)
        out.param_indices.append(i)
        offset, in_bucket = end, in_bucket + p.numel()
        if bucket_size is not None and in_bucket >= bucket_size:
            offset = close(offset)
    if in_bucket > 0 or not out.bucket_indices:
        close(offset)
    return out


def compute_full_layout(params: Iterable[torch.nn.Parameter], data_parallel_world_size: int, expert_data_parallel_world_size: Optional[int] = None,
                        grad_dtype: torch.dtype = torch.float32, bucket_size: Optional[int] = None, sharded: bool = True,
                        pad_for_high_nccl_busbw: bool = False, layer_wise_predicate=None) -> FullParamLayout:
    groups: Dict[BufferKey, List[torch.nn.Parameter]] = {}
    for p in params:
        if not p.requires_grad:
            continue
        storage = torch.uint8 if getattr(p, "is_low_precision_storage", False) else p.dtype
        key = BufferKey(storage, grad_dtype, not getattr(p, "allreduce", True), bool(layer_wise_predicate and layer_wise_predicate(p)))
        groups.setdefault(key, []).append(p)
    full = FullParamLayout()
    for key, ps_ in groups.items():
        dp = (expert_data_parallel_world_size or data_parallel_world_size) if key.is_expert_parallel else data_parallel_world_size
        full.layouts[key] = compute_per_buffer_layout(ps_, dp, bucket_size, sharded, pad_for_high_nccl_busbw, key.is_managed_by_layer_wise_optimizer)
    return full
