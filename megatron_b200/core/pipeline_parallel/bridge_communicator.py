"""Activation exchange between two modules that live on DIFFERENT rank grids (reference ``pipeline_parallel/bridge_communicator.py`` —
``BridgeCommunicator`` :41; used by MIMO models where e.g. a vision encoder runs TP1×DP8 and the language model TP4×DP2).

Both grids are ``HyperCommGrid``s with at least ``tp`` and ``dp`` dims (optionally ``cp``/``pp``).  Only the *boundary* ranks talk: the
last pipeline stage of the source grid and the first stage of the destination grid, and within each data-parallel replica only
the leader (tp rank 0, cp rank 0) — activations are replicated across TP after the row-parallel reduction, so sending every copy
would waste NVLink bandwidth.  The batch dimension is re-partitioned on the way:

* ``src_dp == dst_dp``  — leader i → leader i;
* ``src_dp  > dst_dp``  — fan-in: ``src_dp/dst_dp`` source replicas are concatenated along the batch dim on one destination leader;
* ``src_dp  < dst_dp``  — fan-out: a source leader splits its batch over ``dst_dp/src_dp`` destination leaders.

The destination leader then broadcasts inside its TP×CP group.  ``send_backward`` / ``recv_backward`` run the same plan in reverse
for the gradient.  All point-to-point traffic is posted with ``batch_isend_irecv`` so NCCL can run the pairs concurrently.

The constructor must be called on EVERY rank of the world (process-group creation is collective), including ranks in neither grid.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


class CommRole(Enum):
    SENDER = "sender"          # leader on the source boundary
    RECEIVER = "receiver"      # leader on the destination boundary
    MEMBER = "member"          # non-leader rank of a boundary replica (takes part in the local broadcast only)
    NOOP = "noop"


@dataclass
class RankCommInfo:
    role: CommRole
    peers: List[int]           # global ranks this rank exchanges with, in batch order
    replica: int = -1          # dp index of this rank's replica


def _replica_ranks(grid, boundary: str) -> List[List[int]]:
    """Ranks of each DP replica on the boundary pipeline stage, ordered by dp index; element 0 of a replica is its leader."""
    strides, st = {}, 1
    for name, n in zip(grid.dim_names, grid.shape):      # first dim fastest
        strides[name] = (st, n)
        st *= n
    coord = lambda r, name: (r // strides[name][0]) % strides[name][1] if name in strides else 0  # noqa: E731
    stage = (strides["pp"][1] - 1 if boundary == "last" else 0) if "pp" in strides else 0
    n_dp = strides["dp"][1] if "dp" in strides else 1
    reps: List[List[int]] = [[] for _ in range(n_dp)]
    for r in range(grid.size):
        if coord(r, "pp") == stage:
            reps[coord(r, "dp")].append(r + grid.rank_offset)
    return [sorted(g) for g in reps]


class BridgeCommunicator:
    def __init__(self, src_grid, dst_grid, dim_mapping: Optional[Dict[str, int]] = None, comm_dtype: Optional[torch.dtype] = None):
        self.src_grid, self.dst_grid = src_grid, dst_grid
        self.batch_dim = (dim_mapping or {"s": 0, "b": 1, "h": 2})["b"]
        self.comm_dtype = comm_dtype
        self.rank = dist.get_rank()
        self.src_replicas = _replica_ranks(src_grid, "last")
        self.dst_replicas = _replica_ranks(dst_grid, "first")
        ns, nd = len(self.src_replicas), len(self.dst_replicas)
        if max(ns, nd) % min(ns, nd):
            raise ValueError(f"data-parallel sizes {ns} and {nd} must divide one another")
        # local broadcast groups (collective creation, in a deterministic order)
        self._bcast_group = None
        for reps in (self.src_replicas, self.dst_replicas):
            for rep in reps:
                if len(rep) > 1:
                    pg = dist.new_group(rep)
                    if self.rank in rep:
                        self._bcast_group = pg
        self.src_info = self._info(self.src_replicas, self.dst_replicas, CommRole.SENDER)
        self.dst_info = self._info(self.dst_replicas, self.src_replicas, CommRole.RECEIVER)

    def _info(self, mine: List[List[int]], other: List[List[int]], leader_role: CommRole) -> RankCommInfo:
        for i, rep in enumerate(mine):
            if self.rank in rep:
                n_m, n_o = len(mine), len(other)
                if n_m >= n_o:       # several of mine map to one of theirs
                    peers = [other[i // (n_m // n_o)][0]]
                else:
                    k = n_o // n_m
                    peers = [other[i * k + j][0] for j in range(k)]
                return RankCommInfo(leader_role if self.rank == rep[0] else CommRole.MEMBER, peers, i)
        return RankCommInfo(CommRole.NOOP, [])

    @property
    def is_src(self) -> bool:
        return self.src_info.role is not CommRole.NOOP

    @property
    def is_dst(self) -> bool:
        return self.dst_info.role is not CommRole.NOOP

    # ---- primitives ----
    def _send(self, t: torch.Tensor, info: RankCommInfo) -> None:
        if info.role in (CommRole.MEMBER, CommRole.NOOP):
            return
        t = t.detach()
        if self.comm_dtype is not None:
            t = t.to(self.comm_dtype)
        parts = [t] if len(info.peers) == 1 else list(t.chunk(len(info.peers), dim=self.batch_dim))
        ops = [dist.P2POp(dist.isend, p.contiguous(), peer) for p, peer in zip(parts, info.peers)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def _recv(self, shape: Sequence[int], dtype: torch.dtype, info: RankCommInfo, n_other: int, n_mine: int, device=None) -> torch.Tensor:
        device = device or ("cuda" if torch.cuda.is_available() and dist.get_backend() == "nccl" else "cpu")
        wire = self.comm_dtype or dtype
        out = torch.empty(tuple(shape), dtype=wire, device=device)
        if info.role in (CommRole.SENDER, CommRole.RECEIVER):
            if n_other > n_mine:       # fan-in: my tensor is the concatenation of k peers' tensors
                k = n_other // n_mine
                mine_idx = info.replica
                peers = [p for p in self._peer_leaders(info, k, mine_idx)]
                bufs = [torch.empty_like(c) for c in out.chunk(k, dim=self.batch_dim)]
                ops = [dist.P2POp(dist.irecv, b, peer) for b, peer in zip(bufs, peers)]
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
                out = torch.cat(bufs, dim=self.batch_dim)
            else:                      # 1:1 or fan-out (I receive my slice from one peer)
                for r in dist.batch_isend_irecv([dist.P2POp(dist.irecv, out, info.peers[0])]):
                    r.wait()
        if self._bcast_group is not None and info.role is not CommRole.NOOP:
            leader = dist.get_process_group_ranks(self._bcast_group)[0]
            dist.broadcast(out, src=leader, group=self._bcast_group)
        return out.to(dtype)

    def _peer_leaders(self, info: RankCommInfo, k: int, mine_idx: int) -> List[int]:
        other = self.src_replicas if info is self.dst_info else self.dst_replicas
        return [other[mine_idx * k + j][0] for j in range(k)]

    # ---- public API ----
    def send_forward(self, activation: torch.Tensor) -> None:
        if self.src_info.role is CommRole.SENDER:
            self._send(activation, self.src_info)

    def recv_forward(self, shape: Sequence[int], dtype: torch.dtype = torch.bfloat16, device=None) -> torch.Tensor:
        t = self._recv(shape, dtype, self.dst_info, len(self.src_replicas), len(self.dst_replicas), device)
        return t.requires_grad_(t.is_floating_point())

    def send_backward(self, grad: torch.Tensor) -> None:
        if self.dst_info.role is CommRole.RECEIVER:
            self._send(grad, self.dst_info)

    def recv_backward(self, shape: Sequence[int], dtype: torch.dtype = torch.bfloat16, device=None) -> torch.Tensor:
        return self._recv(shape, dtype, self.src_info, len(self.dst_replicas), len(self.src_replicas), device)

    def send_forward_recv_backward(self, activation: torch.Tensor, grad_shape: Sequence[int], dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        self.send_forward(activation)
        return self.recv_backward(grad_shape, dtype)

    def send_backward_recv_forward(self, grad: torch.Tensor, act_shape: Sequence[int], dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        self.send_backward(grad)
        return self.recv_forward(act_shape, dtype)
