"""Copy-service interface of the refit path (reference ``resharding/copy_services/base.py:13-107``).

A service collects the point-to-point pieces of ONE refit (``submit_send`` / ``submit_recv``) and moves them in ``run()``.
All services here coalesce the pieces of one peer into ONE packed message: on NVLink 5 a refit is launch- and
latency-bound (thousands of KB-sized slices), not bandwidth-bound, so the number of messages is what matters."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class SendOp:
    task_id: Optional[int]
    tensor: torch.Tensor
    dest_rank: int


@dataclass
class RecvOp:
    task_id: Optional[int]
    tensor: torch.Tensor
    src_rank: int


class CopyService(ABC):
    """``submit_*`` only record; ``run()`` executes everything submitted since the last run and clears the queues.
    Ranks are ranks of ``group`` (the joint world of senders and receivers)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.send_ops: List[SendOp] = []
        self.recv_ops: List[RecvOp] = []

    def submit_send(self, src_tensor: torch.Tensor, dest_rank: int, task_id: Optional[int] = None):
        self.send_ops.append(SendOp(task_id, src_tensor, dest_rank))

    def submit_recv(self, dest_tensor: torch.Tensor, src_rank: int, task_id: Optional[int] = None):
        self.recv_ops.append(RecvOp(task_id, dest_tensor, src_rank))

    @abstractmethod
    def run(self):
        ...

    def close(self) -> None:
        self.send_ops.clear()
        self.recv_ops.clear()

    # ---- shared helpers --------------------------------------------------------------------------------------------
    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _take(self) -> Tuple[List[SendOp], List[RecvOp], List[Tuple[SendOp, RecvOp]]]:
        sends, recvs = self.send_ops, self.recv_ops
        self.send_ops, self.recv_ops = [], []
        local, sends, recvs = match_local_ops_by_task_id(sends, recvs, self.rank)
        return sends, recvs, local

    @staticmethod
    def _by_peer(ops, attr: str) -> Dict[int, list]:
        out: Dict[int, list] = {}
        for op in sorted(ops, key=lambda o: (-1 if o.task_id is None else o.task_id)):
            out.setdefault(getattr(op, attr), []).append(op)
        return out


def match_local_ops_by_task_id(send_ops: List[SendOp], recv_ops: List[RecvOp], my_rank: int):
    """Pairs whose source and destination are this rank become plain device copies (reference ``base.py:71``).
    Returns (local pairs, remaining sends, remaining recvs).  Matching is by ``task_id`` when present, else by order."""
    l_send = [s for s in send_ops if s.dest_rank == my_rank]
    l_recv = [r for r in recv_ops if r.src_rank == my_rank]
    if len(l_send) != len(l_recv):
        raise RuntimeError(f"rank {my_rank}: {len(l_send)} local sends but {len(l_recv)} local receives")
    by_id = {r.task_id: r for r in l_recv if r.task_id is not None}
    pairs, rest = [], [r for r in l_recv if r.task_id is None]
    for s in l_send:
        r = by_id.pop(s.task_id, None) if s.task_id is not None else None
        if r is None:
            if not rest:
                raise RuntimeError(f"rank {my_rank}: local send task {s.task_id} has no matching receive")
            r = rest.pop(0)
        pairs.append((s, r))
    return pairs, [s for s in send_ops if s.dest_rank != my_rank], [r for r in recv_ops if r.src_rank != my_rank]


def pack(tensors: List[torch.Tensor], out: Optional[torch.Tensor] = None, align: int = 16) -> Tuple[torch.Tensor, List[int]]:
    """Byte-pack tensors (any dtypes / strides) into one flat uint8 buffer; every piece starts on an ``align`` boundary."""
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += -(-t.numel() * t.element_size() // align) * align
    dev = tensors[0].device if tensors else "cpu"
    buf = out if out is not None else torch.empty(total, dtype=torch.uint8, device=dev)
    assert buf.numel() >= total
    for t, o in zip(tensors, offs):
        nb = t.numel() * t.element_size()
        buf[o:o + nb].view(t.dtype).view(t.shape).copy_(t)
    return buf[:total], offs


def packed_size(shapes_dtypes, align: int = 16) -> int:
    total = 0
    for numel, esize in shapes_dtypes:
        total += -(-numel * esize // align) * align
    return total


def unpack(buf: torch.Tensor, tensors: List[torch.Tensor], wire_dtypes: Optional[List[torch.dtype]] = None, align: int = 16) -> None:
    """Inverse of ``pack``: ``tensors[i]`` (possibly a strided view, possibly another dtype than the wire) is filled."""
    o = 0
    for i, t in enumerate(tensors):
        wd = wire_dtypes[i] if wire_dtypes is not None and wire_dtypes[i] is not None else t.dtype
        nb = t.numel() * torch.empty((), dtype=wd).element_size()
        t.copy_(buf[o:o + nb].view(wd).view(t.shape))
        o += -(-nb // align) * align
