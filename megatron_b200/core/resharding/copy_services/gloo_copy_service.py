"""CPU-staged copy service over a Gloo group (reference ``copy_services/gloo_copy_service.py:15-133``): the fallback for
worlds without NCCL between trainer and server, and what the CPU test-suite runs."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .base import CopyService, pack, packed_size, unpack


class GlooCopyService(CopyService):
    def __init__(self, group=None):
        super().__init__(group)
        self._own_group = None
        if dist.is_initialized() and dist.get_backend(group) != "gloo":
            # a Gloo twin of the (NCCL) group with the same ranks
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
            self._own_group = dist.new_group(ranks=ranks, backend="gloo")
        self._wire_group = self._own_group if self._own_group is not None else group

    def run(self):
        sends, recvs, local = self._take()
        for s, r in local:
            r.tensor.copy_(s.tensor)
        reqs, keep, inbound = [], [], []
        for peer, ops in self._by_peer(sends, "dest_rank").items():
            buf, _ = pack([o.tensor.detach().cpu() if o.tensor.is_cuda else o.tensor.detach() for o in ops])
            keep.append(buf)
            reqs.append(dist.isend(buf, self._global(peer), group=self._wire_group))
        for peer, ops in self._by_peer(recvs, "src_rank").items():
            wire = [getattr(o, "wire_dtype", None) or o.tensor.dtype for o in ops]
            n = packed_size([(o.tensor.numel(), torch.empty((), dtype=w).element_size()) for o, w in zip(ops, wire)])
            buf = torch.empty(n, dtype=torch.uint8)
            inbound.append((buf, ops, wire))
            reqs.append(dist.irecv(buf, self._global(peer), group=self._wire_group))
        for r in reqs:
            r.wait()
        for buf, ops, wire in inbound:
            dev = ops[0].tensor.device
            unpack(buf.to(dev, non_blocking=False) if dev.type != "cpu" else buf, [o.tensor for o in ops], wire)

    def close(self) -> None:
        super().close()
        if self._own_group is not None:
            dist.destroy_process_group(self._own_group)
            self._own_group = None
