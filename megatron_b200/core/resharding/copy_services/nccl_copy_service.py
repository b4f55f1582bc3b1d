"""NCCL copy service (reference ``copy_services/nccl_copy_service.py:15-77``): ONE packed message per peer inside a single
``batch_isend_irecv`` (= one NCCL group launch) on the current stream."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .base import CopyService, pack, packed_size, unpack


class NCCLCopyService(CopyService):
    def run(self):
        sends, recvs, local = self._take()
        for s, r in local:
            r.tensor.copy_(s.tensor)
        p2p, keep, inbound = [], [], []
        for peer, ops in self._by_peer(sends, "dest_rank").items():
            buf, _ = pack([o.tensor.detach() for o in ops])
            keep.append(buf)
            p2p.append(dist.P2POp(dist.isend, buf, self._global(peer), group=self.group))
        for peer, ops in self._by_peer(recvs, "src_rank").items():
            wire = [getattr(o, "wire_dtype", None) or o.tensor.dtype for o in ops]
            n = packed_size([(o.tensor.numel(), torch.empty((), dtype=w).element_size()) for o, w in zip(ops, wire)])
            buf = torch.empty(n, dtype=torch.uint8, device=ops[0].tensor.device)
            inbound.append((buf, ops, wire))
            p2p.append(dist.P2POp(dist.irecv, buf, self._global(peer), group=self.group))
        if p2p:
            for r in dist.batch_isend_irecv(p2p):
                r.wait()
        for buf, ops, wire in inbound:
            unpack(buf, [o.tensor for o in ops], wire)
