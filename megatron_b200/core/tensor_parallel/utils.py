"""Tensor-parallel helpers (reference ``tensor_parallel/utils.py``)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

from .. import parallel_state as ps
from ..utils import divide


def split_tensor_along_last_dim(tensor: torch.Tensor, num_partitions: int, contiguous_split_chunks: bool = False) -> List[torch.Tensor]:
    last = tensor.dim() - 1
    size = divide(tensor.size(last), num_partitions)
    parts = torch.split(tensor, size, dim=last)
    return [p.contiguous() for p in parts] if contiguous_split_chunks else list(parts)


def split_tensor_into_1d_equal_chunks(tensor: torch.Tensor, new_buffer: bool = False, tp_group=None) -> torch.Tensor:
    ws = ps.get_tensor_model_parallel_world_size() if tp_group is None else dist.get_world_size(tp_group)
    rk = ps.get_tensor_model_parallel_rank() if tp_group is None else dist.get_rank(tp_group)
    n = tensor.numel() // ws
    flat = tensor.reshape(-1)[rk * n : (rk + 1) * n]
    return flat.clone() if new_buffer else flat


def gather_split_1d_tensor(tensor: torch.Tensor, tp_group=None) -> torch.Tensor:
    group = ps.get_tensor_model_parallel_group() if tp_group is None else tp_group
    ws = dist.get_world_size(group)
    out = torch.empty(ws * tensor.numel(), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.contiguous(), group=group)
    return out


class VocabUtility:
    """Half-open vocabulary range ``[first, last)`` owned by a TP rank."""

    @staticmethod
    def vocab_range_from_per_partition_vocab_size(per_partition_vocab_size: int, rank: int, world_size: int) -> Sequence[int]:
        first = rank * per_partition_vocab_size
        return first, first + per_partition_vocab_size

    @staticmethod
    def vocab_range_from_global_vocab_size(global_vocab_size: int, rank: int, world_size: int) -> Sequence[int]:
        return VocabUtility.vocab_range_from_per_partition_vocab_size(divide(global_vocab_size, world_size), rank, world_size)
