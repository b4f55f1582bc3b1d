"""Symmetric-memory runtime + NVLink collectives for one process group (``NVLinkBackend``).

Substrate: a per-group *symmetric heap* — one VMM allocation per rank, peer-mapped into every
rank of the group, with an NVLS multicast alias when the fabric supports it.  Allocation and
handle exchange go through ``torch.distributed._symmetric_memory`` (plumbing, like
``torch.distributed`` itself); everything that moves data is our own sm_100a kernel
(``ops/csrc/nvlink_collectives.cu``, ``ops/csrc/fused_tp_gemm.cu``).

Layout of the heap: ``[ flags | workspace A | workspace B | user allocations … ]``.
Workspaces are double-buffered so a collective needs ONE cross-GPU barrier: by the time a
buffer is reused (two ops later) every peer has passed the barrier of the op in between and
therefore finished reading it.

Replaces: NCCL AG/RS/AR on the TP path, TE userbuffers, the NCCL-window allocator
(N3) and the reference's Triton NVLS collectives (§2.4).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops

MAX_RANKS = 16
NUM_SLOTS = 8
_FLAG_BYTES = NUM_SLOTS * MAX_RANKS * 4
_FUSED_FLAG_OFF = 4096  # AG[8*64] + RS[64*8] + XAG[8] uint32 flags of ops/csrc/fused_tp_gemm.cu


def _align(n: int, a: int = 1024) -> int:
    return (n + a - 1) // a * a


class _Handle:
    """Completion handle for collectives issued on the backend's side stream."""

    def __init__(self, event: Optional[torch.cuda.Event]):
        self.event = event

    def wait(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None


from .nvlink_moe import NVLinkMoEMixin


class NVLinkBackend(NVLinkMoEMixin):
    SLOT_MAIN, SLOT_SIDE, SLOT_DDP = 0, 1, 2

    def __init__(self, group, workspace_bytes: Optional[int] = None, heap_bytes: int = 0, use_multicast: Optional[bool] = None, peer_heaps=None):
        """``peer_heaps``: optional list of ``world`` uint8 CUDA tensors (this rank's own allocation at index ``rank``, the others mapped through
        CUDA IPC) to use as the symmetric heap instead of ``torch.distributed._symmetric_memory`` — the P2P (no multicast) protocol then also runs
        between processes that share ONE GPU, which is how the single-GPU CI box exercises the flag protocols (``tests/test_nvlink_ipc_gpu.py``)."""
        assert ops.has_ext() and hasattr(ops.ext(), "nvl_allgather"), "native NVLink kernels are not built"
        if peer_heaps is not None:
            self._init_from_peer_heaps(group, peer_heaps)
            return
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert self.world <= MAX_RANKS
        self.device = torch.device("cuda", torch.cuda.current_device())
        if workspace_bytes is None:
            workspace_bytes = int(os.environ.get("MEGATRON_B200_NVL_WORKSPACE_MB", "320")) << 20
        self.ws_bytes = _align(workspace_bytes, 1 << 16)
        self.heap_bytes = _align(_FLAG_BYTES, 1 << 16) + 2 * self.ws_bytes + _align(heap_bytes, 1 << 16)
        self._symm = symm_mem
        self.buf = symm_mem.empty(self.heap_bytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize()
        try:
            self.hdl = symm_mem.rendezvous(self.buf, group=group)
        except TypeError:
            self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        self.ptrs: List[int] = [int(p) for p in self.hdl.buffer_ptrs]
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if use_multicast is None:
            use_multicast = os.environ.get("MEGATRON_B200_NVLS", "1") != "0"
        self.mc = mc if use_multicast else 0
        self.flags = list(self.ptrs)  # flag arrays live at offset 0 of every heap
        self.ws_off = [_align(_FLAG_BYTES, 1 << 16), _align(_FLAG_BYTES, 1 << 16) + self.ws_bytes]
        self.user_off = self.ws_off[1] + self.ws_bytes
        self._user_cursor = self.user_off
        self.ctrl = torch.zeros(NUM_SLOTS * 4, dtype=torch.int32, device=self.device)
        self.epoch = [0] * NUM_SLOTS
        self._ws_turn = [0] * NUM_SLOTS
        self.side_stream = torch.cuda.Stream()
        self.nblocks = int(os.environ.get("MEGATRON_B200_NVL_BLOCKS", "32"))
        # fused GEMM+collective kernels: chunk flags live in the (otherwise unused) tail of the flag page
        self.fused_flags = [p + _FUSED_FLAG_OFF for p in self.ptrs]
        self.fused_counters = torch.zeros(1024, dtype=torch.int32, device=self.device)
        self.fused_epoch = 0
        # CTA pairs reserved for communication inside a fused kernel (the rest of the 74 run the GEMM): pushes (AG) are
        # fire-and-forget and need fewer; in-switch pull-reductions (RS) are round-trip bound and need more in flight
        self.fused_comm_clusters = [int(os.environ.get("MEGATRON_B200_FUSED_COMM_CLUSTERS_AG", "6")), int(os.environ.get("MEGATRON_B200_FUSED_COMM_CLUSTERS_RS", "10"))]
        self.fused_calls = 0
        dist.barrier(group=group)
        self.barrier()

    def _init_from_peer_heaps(self, group, peer_heaps):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert len(peer_heaps) == self.world and self.world <= MAX_RANKS
        self.device = peer_heaps[self.rank].device
        self._peer_heaps = list(peer_heaps)          # keep the IPC mappings alive
        self.buf = peer_heaps[self.rank]
        self.heap_bytes = self.buf.numel()
        flag_bytes = _align(_FLAG_BYTES, 1 << 16)
        self.ws_bytes = ((self.heap_bytes - flag_bytes) // 2) & ~((1 << 16) - 1)
        self._symm = None
        self.ptrs = [int(t.data_ptr()) for t in peer_heaps]
        self.mc = 0
        self.flags = list(self.ptrs)
        self.ws_off = [flag_bytes, flag_bytes + self.ws_bytes]
        self.user_off = self.ws_off[1] + self.ws_bytes
        self._user_cursor = self.user_off
        self.ctrl = torch.zeros(NUM_SLOTS * 4, dtype=torch.int32, device=self.device)
        self.epoch = [0] * NUM_SLOTS
        self._ws_turn = [0] * NUM_SLOTS
        self.side_stream = torch.cuda.Stream()
        self.nblocks = int(os.environ.get("MEGATRON_B200_NVL_BLOCKS", "32"))
        self.fused_flags = [p + _FUSED_FLAG_OFF for p in self.ptrs]
        self.fused_counters = torch.zeros(1024, dtype=torch.int32, device=self.device)
        self.fused_epoch = 0
        self.fused_comm_clusters = [int(os.environ.get("MEGATRON_B200_FUSED_COMM_CLUSTERS_AG", "6")), int(os.environ.get("MEGATRON_B200_FUSED_COMM_CLUSTERS_RS", "10"))]
        self.fused_calls = 0
        dist.barrier(group=group)
        self.barrier()

    # ---- plumbing -----------------------------------------------------------------------------------
    def _next_epoch(self, slot: int, n: int = 1) -> int:
        e = self.epoch[slot] + 1
        self.epoch[slot] += n
        return e

    def _view(self, off: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.buf[off : off + nbytes].view(dtype)

    def _workspace(self, slot: int, nbytes: int) -> int:
        """Alternate between the two workspaces; main and side slots use disjoint halves."""
        assert nbytes <= self.ws_bytes // 2, f"NVLink workspace too small: need {nbytes} bytes, have {self.ws_bytes // 2} (MEGATRON_B200_NVL_WORKSPACE_MB)"
        turn = self._ws_turn[slot]
        self._ws_turn[slot] ^= 1
        half = 0 if slot == self.SLOT_MAIN else self.ws_bytes // 2
        return self.ws_off[turn] + half

    def alloc_symmetric(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        """Bump-allocate from the user region (all ranks must call in the same order)."""
        nbytes = _align(numel * torch.empty((), dtype=dtype).element_size(), 1 << 16)
        if self._user_cursor + nbytes > self.heap_bytes:
            # grow: separate symmetric allocation with its own peer mapping
            extra = _ExtraRegion(self, nbytes)
            self._extras = getattr(self, "_extras", []) + [extra]
            return extra.tensor(numel, dtype)
        off = self._user_cursor
        self._user_cursor += nbytes
        t = self._view(off, numel, dtype)
        t.zero_()
        return t

    def _region_of(self, t: torch.Tensor) -> Optional[Tuple[List[int], int, int]]:
        """(peer base pointers, multicast base, byte offset) if ``t`` lives in symmetric memory."""
        p = t.data_ptr()
        if self.ptrs[self.rank] <= p < self.ptrs[self.rank] + self.heap_bytes:
            return self.ptrs, self.mc, p - self.ptrs[self.rank]
        for ex in getattr(self, "_extras", []):
            if ex.ptrs[self.rank] <= p < ex.ptrs[self.rank] + ex.nbytes:
                return ex.ptrs, ex.mc, p - ex.ptrs[self.rank]
        return None

    def barrier(self, slot: int = 0):
        ops.ext().nvl_barrier(self.ptrs, self.flags, self.rank, self._next_epoch(slot), slot)

    # ---- raw collectives -------------------------------------------------------------------------------
    def _ag_into(self, x: torch.Tensor, ptrs, mc, dst_off: int, slot: int):
        ops.ext().nvl_allgather(ptrs, self.flags, mc, x, dst_off, self.rank, self._next_epoch(slot), self.ctrl, slot, self.nblocks)
        ops._count()

    def all_gather(self, x: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """[n, …] → [world*n, …] (view into the symmetric workspace; valid until two collectives later)."""
        x = x.contiguous()
        nbytes = x.numel() * x.element_size()
        off = self._workspace(slot, nbytes * self.world)
        self._ag_into(x, self.ptrs, self.mc, off, slot)
        return self._view(off, x.numel() * self.world, x.dtype).view(x.shape[0] * self.world, *x.shape[1:])

    def symmetric_like(self, shape, dtype, slot: int = 0) -> torch.Tensor:
        """Workspace tensor a producer kernel (GEMM epilogue) can write so that peers can read it."""
        numel = 1
        for s in shape:
            numel *= s
        off = self._workspace(slot, numel * torch.empty((), dtype=dtype).element_size())
        return self._view(off, numel, dtype).view(*shape)

    def reduce_scatter(self, x: torch.Tensor, slot: int = 0, scale: float = 1.0, out: Optional[torch.Tensor] = None, trailing: bool = False) -> torch.Tensor:
        """[world*n, …] → [n, …], summed over ranks with fp32 accumulation in the switch."""
        reg = self._region_of(x)
        if reg is None:
            ws = self.symmetric_like(x.shape, x.dtype, slot)
            ws.copy_(x)
            x, reg = ws, self._region_of(ws)
        ptrs, mc, off = reg
        n0 = x.shape[0] // self.world
        if out is None:
            out = torch.empty((n0, *x.shape[1:]), dtype=x.dtype, device=x.device)
        e = self._next_epoch(slot, 2)
        ops.ext().nvl_reducescatter(ptrs, self.flags, mc, off, out, float(scale), self.rank, e, self.ctrl, slot, trailing, self.nblocks)
        ops._count()
        return out

    def all_reduce(self, x: torch.Tensor, slot: int = 0, scale: float = 1.0) -> torch.Tensor:
        reg = self._region_of(x)
        copy_back = None
        if reg is None:
            ws = self.symmetric_like(x.shape, x.dtype, slot)
            ws.copy_(x)
            copy_back, x, reg = x, ws, self._region_of(ws)
        ptrs, mc, off = reg
        code = {torch.float32: 0, torch.bfloat16: 1}[x.dtype]
        e = self._next_epoch(slot, 2)
        ops.ext().nvl_allreduce(ptrs, self.flags, mc, off, x.numel(), code, float(scale), self.rank, e, self.ctrl, slot, self.nblocks)
        ops._count()
        if copy_back is not None:
            copy_back.copy_(x)
            return copy_back
        return x

    def all_reduce_async(self, t: torch.Tensor) -> _Handle:
        s = self.side_stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.all_reduce(t, slot=self.SLOT_SIDE)
            ev = torch.cuda.Event()
            ev.record(s)
        t.record_stream(s)
        return _Handle(ev)

    # ---- DDP / distributed-optimizer buffers (must be allocated with alloc_symmetric) --------------------
    def _run_ddp(self, fn, async_op: bool, *keep):
        """Run a DDP collective.  ``async_op``: on this backend's side stream, ordered after the work already queued on the caller's stream; returns a handle
        whose ``wait()`` makes the caller's stream wait for it (reference semantics of ``async_op=True``: param_and_grad_buffer.py:600-785 — the bucket's
        reduce-scatter overlaps the rest of the backward pass, the parameter all-gather overlaps the next forward).  Ops of one backend share the side
        stream and the DDP flag slot, so they execute in issue order on every rank."""
        if not async_op:
            fn()
            return None
        cur = torch.cuda.current_stream()
        s = self.side_stream
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            fn()
            ev = torch.cuda.Event()
            ev.record(s)
        for t in keep:
            t.record_stream(s)
        return _Handle(ev)

    def reduce_scatter_scaled_(self, grad_data: torch.Tensor, scale: float, async_op: bool = False):
        """In place: my shard of ``grad_data`` <- scale * sum over ranks of that shard (scale + reduce + cast fused in one multimem kernel)."""
        n = grad_data.numel() // self.world
        mine = grad_data[self.rank * n : (self.rank + 1) * n]
        if self._region_of(grad_data) is None:
            return dist.reduce_scatter_tensor(mine, grad_data.mul_(scale), group=self.group, async_op=async_op)
        return self._run_ddp(lambda: self.reduce_scatter(grad_data.view(self.world, n), slot=self.SLOT_DDP, scale=scale, out=mine.view(1, n), trailing=True), async_op, grad_data)

    def all_reduce_scaled_(self, grad_data: torch.Tensor, scale: float, async_op: bool = False):
        if self._region_of(grad_data) is None:
            return dist.all_reduce(grad_data.mul_(scale), group=self.group, async_op=async_op)
        return self._run_ddp(lambda: self.all_reduce(grad_data, slot=self.SLOT_DDP, scale=scale), async_op, grad_data)

    def all_gather_inplace_(self, param_data: torch.Tensor, async_op: bool = False):
        """Publish my shard of ``param_data`` to the same offset on every rank (multicast store)."""
        n = param_data.numel() // self.world
        mine = param_data[self.rank * n : (self.rank + 1) * n]
        reg = self._region_of(param_data)
        if reg is None:
            return dist.all_gather_into_tensor(param_data, mine.clone(), group=self.group, async_op=async_op)
        ptrs, mc, off = reg
        return self._run_ddp(lambda: self._ag_into(mine, ptrs, mc, off, self.SLOT_DDP), async_op, param_data)

    # ---- pair ops used by the TP layers ---------------------------------------------------------------------
    def _fused_ok(self, M: int, N: int, K: int, *ts) -> bool:
        from . import fused

        if fused.get_mode(world_size=self.world) != "fused" or not hasattr(ops.ext(), "fused_tp_gemm"):
            return False
        if self.world > 8 or M % (self.world * 256) != 0 or M // self.world // 256 > 64 or K % 8 != 0 or N % 8 != 0:
            return False
        return all(t.dtype == torch.bfloat16 for t in ts)

    def _fused_launch(self, mode, a, b, c, b_layout, ag_src=None, ag_off=0, rs_off=0, rs_out=None, xag_src=None, xag_off=0):
        self.fused_epoch += 1
        empty = a.new_empty(0)
        mc = self.mc
        ops.ext().fused_tp_gemm(
            mode, a, b, c, b_layout, self.rank, self.fused_epoch,
            ag_src if ag_src is not None else empty, (mc + ag_off) if (mc and mode == 0) else 0, [p + ag_off for p in self.ptrs] if mode == 0 else [],
            (mc + rs_off) if (mc and mode >= 1) else 0, [p + rs_off for p in self.ptrs] if mode >= 1 else [], rs_out if rs_out is not None else empty,
            xag_src if xag_src is not None else empty, (mc + xag_off) if (mc and xag_src is not None) else 0,
            [p + xag_off for p in self.ptrs] if xag_src is not None else [], self.fused_flags, self.fused_counters, self.fused_comm_clusters[min(mode, 1)],
        )
        self.fused_calls += 1
        ops._count()

    def _fused_ag_gemm(self, x: torch.Tensor, w: torch.Tensor, b_layout: int):
        """One kernel: multicast-push my shard of ``x`` chunk by chunk while tcgen05 CTAs consume arrived chunks.
        Returns (out [M, N], gathered x [M, K] living in the symmetric workspace)."""
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        K = x2.shape[1]
        M = x2.shape[0] * self.world
        N = w.shape[0] if b_layout == 0 else w.shape[1]
        off = self._workspace(self.SLOT_MAIN, M * K * 2)
        full = self._view(off, M * K, x.dtype).view(M, K)
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
        self._fused_launch(0, full, w, out, b_layout, ag_src=x2, ag_off=off)
        return out, full

    def _fused_gemm_rs(self, x: torch.Tensor, w: torch.Tensor, b_layout: int, xag: Optional[torch.Tensor] = None):
        """One kernel: tcgen05 CTAs write partial tiles to symmetric memory, signal the owner rank per 256-row
        block; reducer CTAs pull-reduce through the switch (``multimem.ld_reduce``).  Optionally the reducer
        CTAs also all-gather ``xag`` (the wgrad operand) first.  Returns (out [M/world, N], gathered xag)."""
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        M, K = x2.shape
        N = w.shape[0] if b_layout == 0 else w.shape[1]
        ybytes = _align(M * N * 2, 1 << 12)
        xbytes = 0
        if xag is not None:
            xag = xag.reshape(-1, xag.shape[-1]).contiguous()
            xbytes = xag.numel() * xag.element_size() * self.world
        off = self._workspace(self.SLOT_MAIN, ybytes + xbytes)
        y = self._view(off, M * N, x.dtype).view(M, N)
        out = torch.empty((M // self.world, N), dtype=x.dtype, device=x.device)
        full = None
        if xag is not None:
            full = self._view(off + ybytes, xag.numel() * self.world, xag.dtype).view(xag.shape[0] * self.world, xag.shape[1])
        self._fused_launch(1, x2, w, y, b_layout, rs_off=off, rs_out=out, xag_src=xag, xag_off=off + ybytes)
        return out, full

    def _fused_gemm_ar(self, x: torch.Tensor, w: torch.Tensor, b_layout: int) -> torch.Tensor:
        """One kernel: tcgen05 CTAs write partial tiles to symmetric memory; the owner of each 256-row block reduces it in the
        switch (``multimem.ld_reduce``) and broadcasts the sum back into EVERY rank's buffer (``multimem.st``), in place.
        Returns a VIEW of the symmetric workspace [M, N] (valid until two collectives later: consume or copy at once)."""
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        M, K = x2.shape
        N = w.shape[0] if b_layout == 0 else w.shape[1]
        off = self._workspace(self.SLOT_MAIN, _align(M * N * 2, 1 << 12))
        y = self._view(off, M * N, x.dtype).view(M, N)
        self._fused_launch(2, x2, w, y, b_layout, rs_off=off)
        return y

    def gemm_all_reduce(self, x: torch.Tensor, w: torch.Tensor, b_layout: int = 0) -> torch.Tensor:
        """``X op(W)`` summed over the group (non-SP row-parallel forward: b_layout 0 = W[N,K]; column-parallel dgrad: 1 = W[K,N])."""
        rows = x.numel() // x.shape[-1]
        N = w.shape[0] if b_layout == 0 else w.shape[1]
        if self._fused_ok(rows, N, x.shape[-1], x, w):
            return self._fused_gemm_ar(x, w, b_layout).view(*x.shape[:-1], N).clone()
        y = self.symmetric_like((*x.shape[:-1], N), x.dtype)
        (ops.gemm_nt if b_layout == 0 else ops.gemm_nn)(x, w, out=y)
        return self.all_reduce(y).clone()

    def all_gather_gemm(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        rows = x.numel() // x.shape[-1]
        if self._fused_ok(rows * self.world, w.shape[0], x.shape[-1], x, w):
            out, _ = self._fused_ag_gemm(x, w, 0)
            return out.view(x.shape[0] * self.world, *x.shape[1:-1], w.shape[0])
        full = self.all_gather(x)
        return ops.gemm_nt(full, w)

    def gemm_reduce_scatter(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        rows = x.numel() // x.shape[-1]
        if self._fused_ok(rows, w.shape[0], x.shape[-1], x, w):
            out, _ = self._fused_gemm_rs(x, w, 0)
            return out.view(x.shape[0] // self.world, *x.shape[1:-1], w.shape[0])
        y = self.symmetric_like((*x.shape[:-1], w.shape[0]), x.dtype)
        ops.gemm_nt(x, w, out=y.view(-1, w.shape[0]))
        return self.reduce_scatter(y)

    def sp_linear_backward(self, gy, x, weight, wgrad_needed: bool, accumulate: bool, wgrad_fn):
        """Column-parallel backward under SP.  Fused: ONE kernel does dgrad GEMM → reduce-scatter AND the
        all-gather of ``x`` for wgrad.  Unfused: dgrad GEMM → RS on the main stream, AG(x) on the side stream."""
        rows = gy.numel() // gy.shape[-1]
        if self._fused_ok(rows, weight.shape[1], gy.shape[-1], gy, weight, x):
            gx, full_x = self._fused_gemm_rs(gy, weight, 1, xag=x if wgrad_needed else None)
            gx = gx.view(gy.shape[0] // self.world, *gy.shape[1:-1], weight.shape[1])
            gw = wgrad_fn(gy, full_x, weight, accumulate) if wgrad_needed else None
            return gx, gw
        cur = torch.cuda.current_stream()
        full_x = ev = None
        if wgrad_needed:
            s = self.side_stream
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                full_x = self.all_gather(x, slot=self.SLOT_SIDE)
                ev = torch.cuda.Event()
                ev.record(s)
        gx_full = self.symmetric_like((*gy.shape[:-1], weight.shape[1]), gy.dtype)
        ops.gemm_nn(gy, weight, out=gx_full)
        gx = self.reduce_scatter(gx_full)
        gw = None
        if wgrad_needed:
            cur.wait_event(ev)
            gw = wgrad_fn(gy, full_x, weight, accumulate)
        return gx, gw

    def row_linear_backward_sp(self, gy, x, weight, wgrad_needed: bool, accumulate: bool, wgrad_fn):
        rows = gy.numel() // gy.shape[-1]
        if self._fused_ok(rows * self.world, weight.shape[1], gy.shape[-1], gy, weight):
            gx, full_gy = self._fused_ag_gemm(gy, weight, 1)
            gx = gx.view(gy.shape[0] * self.world, *gy.shape[1:-1], weight.shape[1])
            gw = wgrad_fn(full_gy, x, weight, accumulate) if wgrad_needed else None
            return gx, gw
        full_gy = self.all_gather(gy)
        gx = ops.gemm_nn(full_gy, weight)
        gw = wgrad_fn(full_gy, x, weight, accumulate) if wgrad_needed else None
        return gx, gw


class _ExtraRegion:
    """A separately rendezvoused symmetric allocation (DDP buffers larger than the heap)."""

    def __init__(self, be: NVLinkBackend, nbytes: int):
        self.nbytes = nbytes
        self.buf = be._symm.empty(nbytes, dtype=torch.uint8, device=be.device)
        self.buf.zero_()
        torch.cuda.synchronize()
        try:
            self.hdl = be._symm.rendezvous(self.buf, group=be.group)
        except TypeError:
            self.hdl = be._symm.rendezvous(self.buf, be.group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.mc = mc if be.mc else 0

    def tensor(self, numel, dtype):
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.buf[:nbytes].view(dtype)
