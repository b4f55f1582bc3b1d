"""Unevenly sharded flat tensors (reference ``megatron_fsdp/uneven_dtensor.py:1-483``).

An FSDP shard boundary falls wherever ``bucket_size / world`` puts it, so a parameter is generally split UNEVENLY: rank 0 may
own 3 rows and a fraction, rank 1 the rest, rank 2 nothing.  ``torch DTensor`` assumes even ``Shard(0)`` chunks; the reference
carries explicit chunk metadata on the DTensor.  Here the same information is a small dataclass next to a plain local tensor —
(global shape, [start, end) of the local FLAT slice) — which is all checkpointing (1-D ranged ``ShardedTensor``, see
``transformer/fsdp_dtensor_checkpoint.py``) and full-tensor reconstruction need."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass
class UnevenShard:
    """This rank's flat slice ``[start, end)`` of a tensor of ``global_shape`` (may be empty)."""
    local: torch.Tensor
    global_shape: Tuple[int, ...]
    start: int
    end: int

    @property
    def numel(self) -> int:
        n = 1
        for d in self.global_shape:
            n *= d
        return n


def split_dtensor(full: torch.Tensor, boundaries: Sequence[int], rank: int) -> UnevenShard:
    """Slice ``full`` (flattened) at ``boundaries`` (world+1 ascending offsets, not necessarily uniform)."""
    s, e = int(boundaries[rank]), int(boundaries[rank + 1])
    return UnevenShard(full.reshape(-1)[s:e].clone(), tuple(full.shape), s, e)


def gather_and_compute_chunk_metadata(shard: UnevenShard, group=None) -> List[Tuple[int, int]]:
    """Every rank's [start, end) (reference :gather_and_compute_chunk_metadata)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [(shard.start, shard.end)]
    out: List = [None] * world
    dist.all_gather_object(out, (shard.start, shard.end), group=group)
    return out


def update_uneven_dtensor_chunk_metadata(shard: UnevenShard, start: int, end: int) -> UnevenShard:
    assert end - start == shard.local.numel(), "chunk metadata must describe the local tensor"
    shard.start, shard.end = start, end
    return shard


def validate_uneven_dtensor(shard: UnevenShard, group=None) -> None:
    """The chunks of all ranks must tile [0, numel) exactly once."""
    chunks = sorted(c for c in gather_and_compute_chunk_metadata(shard, group) if c[1] > c[0])
    pos = 0
    for s, e in chunks:
        if s != pos:
            raise ValueError(f"uneven shards {'overlap' if s < pos else 'leave a hole'} at flat offset {min(s, pos)}: {chunks}")
        pos = e
    if pos != shard.numel:
        raise ValueError(f"uneven shards cover {pos} of {shard.numel} elements: {chunks}")


def gather_uneven_dtensor_to_full_tensor(shard: UnevenShard, group=None) -> torch.Tensor:
    """All ranks obtain the full tensor: pad every chunk to the largest, ONE all-gather, un-pad (NCCL has no all-gather-v)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return shard.local.view(shard.global_shape).clone()
    chunks = gather_and_compute_chunk_metadata(shard, group)
    m = max(e - s for s, e in chunks)
    pad = torch.zeros(m, dtype=shard.local.dtype, device=shard.local.device)
    pad[: shard.local.numel()].copy_(shard.local)
    out = torch.empty(world * m, dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.empty(shard.numel, dtype=pad.dtype, device=pad.device)
    for r, (s, e) in enumerate(chunks):
        full[s:e].copy_(out[r * m: r * m + (e - s)])
    return full.view(shard.global_shape)


uneven_dtensor_to_full_tensor = gather_uneven_dtensor_to_full_tensor


def redistribute_uneven_dtensor_to_replicated(shard: UnevenShard, group=None) -> UnevenShard:
    full = gather_uneven_dtensor_to_full_tensor(shard, group)
    return UnevenShard(full.reshape(-1), shard.global_shape, 0, full.numel())


def get_unflattened_state_dict(flat: Dict[str, object], sep: str = ".") -> Dict:
    out: Dict = {}
    for k, v in flat.items():
        d = out
        parts = k.split(sep)
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def filter_unflattened_state_dict(sd: Dict, key_chain: Sequence[str] = (), visit_condition=lambda v: isinstance(v, UnevenShard)) -> List[Tuple[Tuple[str, ...], object]]:
    out = []
    for k, v in sd.items():
        if isinstance(v, dict):
            out.extend(filter_unflattened_state_dict(v, tuple(key_chain) + (k,), visit_condition))
        elif visit_condition(v):
            out.append((tuple(key_chain) + (k,), v))
    return out


def preprocess_state_dict_for_uneven_dtensor(state_dict: Dict, group=None) -> Dict:
    """Validate every uneven shard of a (nested) state dict before it is handed to the checkpoint writer."""
    for _, v in filter_unflattened_state_dict(state_dict):
        validate_uneven_dtensor(v, group)
    return state_dict
