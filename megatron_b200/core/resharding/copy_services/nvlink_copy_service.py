"""Refit over NVLink peer memory — this framework's replacement for the reference's NVSHMEM copy service
(``copy_services/nvshmem_copy_service.py:1-166`` + ``nvshmem_copy_service/`` — SURVEY N2).

Every rank owns a receive slab in the symmetric heap (``parallel/nvlink.py``).  ``run()``:

1. (first run of a plan) the ranks exchange, in one ``all_gather_object``, how many bytes they send to every peer; each receiver
   lays the inbound regions out back to back, so each sender knows the byte offset of its region in every peer's slab — cached
   per signature, so steady-state refits do no host collectives at all;
2. ONE ``batched_copy`` kernel per rank stores every contiguous piece straight into the peers' slabs through their mapped
   pointers (``ops/csrc/runtime_native.cu::batched_copy_kernel``: 64 KiB chunks spread over the grid, 16-byte vector
   stores); strided pieces are first made contiguous into a local staging buffer by PyTorch;
3. one cross-GPU barrier (``nvl_barrier``: release/acquire flags in the heap);
4. the receiver scatters its slab into the destination slices (dtype conversion included).

No NVSHMEM, no NCCL: the transfer is plain stores over NVSwitch issued by SMs, the same mechanism as the MoE dispatcher."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .base import CopyService, packed_size, unpack


class NVLinkCopyService(CopyService):
    ALIGN = 16

    def __init__(self, group=None, backend=None, slab_bytes: int = 0):
        super().__init__(group)
        if backend is None:
            from ....parallel.collectives import enable_for_group
            backend = enable_for_group(group)         # collective: creates (or reuses) the symmetric heap of this group
        self.be = backend
        self._slab: Optional[torch.Tensor] = None
        self._slab_bytes = slab_bytes
        self._layouts: Dict[Tuple, Tuple[Dict[int, int], Dict[int, int]]] = {}

    def _ensure_slab(self, nbytes: int):
        # symmetric allocation: every rank must allocate the same size in the same order -> agree on the maximum
        if self._slab is not None and self._slab.numel() >= nbytes:
            return
        need = torch.tensor([nbytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        self._slab = self.be.alloc_symmetric(max(int(need.item()), self._slab_bytes, 1 << 20), torch.uint8)

    def _layout(self, send_sizes: Dict[int, int], recv_sizes: Dict[int, int]):
        sig = (tuple(sorted(send_sizes.items())), tuple(sorted(recv_sizes.items())))
        hit = self._layouts.get(sig)
        if hit is not None:
            return hit
        # my inbound regions, ordered by sender
        base, cur = {}, 0
        for src in sorted(recv_sizes):
            base[src] = cur
            cur += recv_sizes[src]
        gathered: List = [None] * self.world_size
        dist.all_gather_object(gathered, (base, cur), group=self.group)
        self._ensure_slab(max(g[1] for g in gathered))
        remote = {dst: gathered[dst][0][self.rank] for dst in send_sizes}
        self._layouts[sig] = (base, remote)
        return base, remote

    def run(self):
        from .... import ops
        sends, recvs, local = self._take()
        for s, r in local:
            r.tensor.copy_(s.tensor)
        s_by, r_by = self._by_peer(sends, "dest_rank"), self._by_peer(recvs, "src_rank")
        wire = {p: [getattr(o, "wire_dtype", None) or o.tensor.dtype for o in ops_] for p, ops_ in r_by.items()}
        send_sizes = {p: packed_size([(o.tensor.numel(), o.tensor.element_size()) for o in ops_], self.ALIGN) for p, ops_ in s_by.items()}
        recv_sizes = {p: packed_size([(o.tensor.numel(), torch.empty((), dtype=w).element_size()) for o, w in zip(ops_, wire[p])], self.ALIGN) for p, ops_ in r_by.items()}
        base, remote = self._layout(send_sizes, recv_sizes)
        region = self.be._region_of(self._slab)
        assert region is not None, "the receive slab must live in symmetric memory"
        peer_ptrs, _, slab_off = region
        tasks, keep = [], []
        for peer, ops_ in s_by.items():
            off = remote[peer]
            for o in ops_:
                t = o.tensor.detach()
                if not t.is_contiguous():
                    t = t.contiguous()
                    keep.append(t)
                nb = t.numel() * t.element_size()
                tasks.append((t.data_ptr(), peer_ptrs[peer] + slab_off + off, nb))
                off += -(-nb // self.ALIGN) * self.ALIGN
        self.be.barrier()                                 # nobody may still be reading its slab from the previous refit
        if tasks:
            ops.ext().batched_copy(torch.tensor(tasks, dtype=torch.int64), 148 * 8)   # 8 CTAs per SM keep enough 16-byte stores in flight
        self.be.barrier()                                 # stores visible at the receivers
        for peer, ops_ in r_by.items():
            unpack(self._slab[base[peer]: base[peer] + recv_sizes[peer]], [o.tensor for o in ops_], wire[peer], self.ALIGN)
        del keep
