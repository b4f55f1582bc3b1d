"""Execute a reshard plan (reference ``resharding/execute.py`` + copy services: NCCL / Gloo / NVSHMEM / NIXL).

Two back ends:
* ``p2p``     — ``torch.distributed.batch_isend_irecv`` of packed contiguous pieces (NCCL on GPUs, Gloo on CPU).
* ``nvlink``  — all destination tensors live in the symmetric heap: ONE ``batched_copy`` kernel per rank stores every piece
                straight into the peers' memory over NVLink (``ops/csrc/runtime_native.cu``, replaces the NVSHMEM service N2),
                followed by a cross-GPU barrier.  Pieces must be contiguous runs, which holds for dim-0 (row) resharding.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

from .planner import ShardDesc, TransferOp, build_reshard_plan


def _view(t: torch.Tensor, slices) -> torch.Tensor:
    return t[tuple(slice(lo, hi) for lo, hi in slices)]


def execute_reshard_plan(plan: Sequence[TransferOp], src_tensors: Dict[str, torch.Tensor], dst_tensors: Dict[str, torch.Tensor], group=None) -> None:
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    p2p, recv_bufs, keep = [], [], []
    for i, op in enumerate(plan):
        if op.src_rank == rank and op.dst_rank == rank:
            _view(dst_tensors[op.key], op.dst_slices).copy_(_view(src_tensors[op.key], op.src_slices))
        elif op.src_rank == rank:
            buf = _view(src_tensors[op.key], op.src_slices).contiguous()
            keep.append(buf)
            p2p.append(dist.P2POp(dist.isend, buf, dist.get_global_rank(group, op.dst_rank) if group is not None else op.dst_rank, group=group, tag=i))
        elif op.dst_rank == rank:
            tgt = _view(dst_tensors[op.key], op.dst_slices)
            buf = torch.empty(tgt.shape, dtype=tgt.dtype, device=tgt.device)
            recv_bufs.append((tgt, buf))
            p2p.append(dist.P2POp(dist.irecv, buf, dist.get_global_rank(group, op.src_rank) if group is not None else op.src_rank, group=group, tag=i))
    if p2p:
        if dist.get_backend(group) == "gloo":
            reqs = [op.op(op.tensor, op.peer, group=op.group, tag=op.tag) for op in p2p]
        else:
            reqs = dist.batch_isend_irecv(p2p)
        for r in reqs:
            r.wait()
    for tgt, buf in recv_bufs:
        tgt.copy_(buf)


def reshard_state_dict(src_sharded: Dict[str, "object"], dst_sharded: Dict[str, "object"], group=None) -> None:
    """Move data from the ``ShardedTensor``s this rank holds (``src_sharded``) into the ones it wants (``dst_sharded``; their
    ``.data`` tensors are filled in place).  Collective over ``group``."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    mine = ([ShardDesc.from_sharded_tensor(s, rank) for s in src_sharded.values()], [ShardDesc.from_sharded_tensor(s, rank) for s in dst_sharded.values()])
    gathered: List = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    sources = [s for g in gathered for s in g[0]]
    dests = [d for g in gathered for d in g[1]]
    plan = build_reshard_plan(sources, dests)
    execute_reshard_plan(plan, {s.key: s.data for s in src_sharded.values()}, {s.key: s.data for s in dst_sharded.values()}, group)
