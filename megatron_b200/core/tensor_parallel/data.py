"""Broadcast a batch from TP-rank 0 to its TP group (reference ``tensor_parallel/data.py:64``)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from .. import parallel_state as ps

_MAX_DATA_DIM = 5


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and dist.get_backend() != "gloo" else torch.device("cpu")


def _build_key_size_numel_dictionaries(keys: List[str], data, tp_group=None):
    group = tp_group or ps.get_tensor_model_parallel_group()
    src = dist.get_process_group_ranks(group)[0]
    sizes = [0] * (_MAX_DATA_DIM * len(keys))
    if dist.get_rank() == src:
        for i, k in enumerate(keys):
            assert data[k].dim() < _MAX_DATA_DIM
            for j, s in enumerate(data[k].shape):
                sizes[i * _MAX_DATA_DIM + j] = s
    t = torch.tensor(sizes, dtype=torch.long, device=_dev())
    dist.broadcast(t, src, group=group)
    sizes = t.cpu().tolist()
    key_size, key_numel, total = {}, {}, 0
    for i, k in enumerate(keys):
        shape = [s for s in sizes[i * _MAX_DATA_DIM : (i + 1) * _MAX_DATA_DIM] if s > 0]
        n = 1
        for s in shape:
            n *= s
        key_size[k], key_numel[k] = shape, n
        total += n
    return key_size, key_numel, total


def broadcast_data(keys: List[str], data: Dict[str, torch.Tensor], datatype, tp_group=None) -> Dict[str, torch.Tensor]:
    """All tensors under ``keys`` are flattened into ONE buffer and broadcast once."""
    group = tp_group or ps.get_tensor_model_parallel_group()
    src = dist.get_process_group_ranks(group)[0]
    key_size, key_numel, total = _build_key_size_numel_dictionaries(keys, data, group)
    if dist.get_rank() == src:
        for k in keys:
            assert data[k].dtype == datatype, f"{k} has dtype {data[k].dtype}, expected {datatype}"
        flat = torch.cat([data[k].contiguous().view(-1) for k in keys]).to(_dev())
    else:
        flat = torch.empty(total, dtype=datatype, device=_dev())
    dist.broadcast(flat, src, group=group)
    out, off = {}, 0
    for k in keys:
        out[k] = flat[off : off + key_numel[k]].view(key_size[k])
        off += key_numel[k]
    return out
