"""Expert-parallel dispatch / combine over the symmetric heap (``MoEFlexTokenDispatcher``'s engine).

dispatch = ONE push kernel: every (token, expert) pair's row is stored straight into the destination rank's receive buffer at the
slot that makes the buffer grouped by local expert (then by source rank) — the all-to-all and both permutations of the reference's
A2A dispatcher (``token_dispatcher.py:375-960``) collapse into one pass over NVLink.  combine = ONE pull kernel: each token reads
its k expert outputs from the peers' buffers and sums them in fp32.  Slot arithmetic uses a tiny all-gather of the per-expert
counts ([W, E] int32); the only host sync is the receive-size read-back the grouped GEMM needs anyway.

Autograd: push and pull are adjoint (dispatch.backward = pull of the gradients, combine.backward = push), so four functions cover
forward and backward of both directions."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .. import ops


@dataclass
class DispatchHandle:
    T: int
    topk: int
    src_tok: torch.Tensor        # [P] source token of each pair (expert-major order)
    dst_rank: torch.Tensor       # [P] int32
    dst_slot: torch.Tensor       # [P] int64 row in the destination buffer
    tok_rank: torch.Tensor       # [T, topk] int32   per-token view of the same pairs (−1 padded slots)
    tok_slot: torch.Tensor       # [T, topk] int64
    m_recv: int                  # rows this rank receives


class NVLinkMoEMixin:
    """Methods added to ``NVLinkBackend`` (needs: ptrs, world, rank, _workspace, _view, barrier, all_gather, SLOT_MAIN)."""

    # ---- primitives ---------------------------------------------------------------------------------------------------
    def _moe_push(self, x: torch.Tensor, src_row, dst_rank, dst_slot, m_recv: int) -> torch.Tensor:
        """Rows of ``x`` → peers' receive buffers; returns this rank's received rows [m_recv, cols] (a private copy)."""
        x = x.contiguous()
        row_bytes = x.shape[1] * x.element_size()
        cap = self.ws_bytes // 2
        assert m_recv * row_bytes <= cap, f"MoE receive buffer needs {m_recv * row_bytes} bytes > workspace {cap} (MEGATRON_B200_NVL_WORKSPACE_MB)"
        off = self._workspace(self.SLOT_MAIN, max(16, m_recv * row_bytes))
        ops.ext().moe_push_rows(x, src_row, dst_rank, dst_slot, self.ptrs, off)
        ops._count()
        self.barrier()  # every peer's pushes have landed here
        return self._view(off, m_recv * x.shape[1], x.dtype).view(m_recv, x.shape[1]).clone()

    def _moe_pull(self, y: torch.Tensor, rank_tbl, slot_tbl, w: Optional[torch.Tensor], raw: bool) -> torch.Tensor:
        """Publish ``y`` (rows grouped as received) and gather/sum the rows each output element needs from the peers."""
        y = y.contiguous()
        nbytes = y.numel() * y.element_size()
        off = self._workspace(self.SLOT_MAIN, max(16, nbytes))
        if y.numel():
            self._view(off, y.numel(), y.dtype).view_as(y).copy_(y)
        self.barrier()  # all ranks' buffers are complete
        out = ops.ext().moe_pull_rows(rank_tbl, slot_tbl, None if w is None else w.float().contiguous(), self.ptrs, off, y.shape[1], y.dtype, raw)
        ops._count()
        return out

    # ---- public API used by MoEFlexTokenDispatcher ----------------------------------------------------------------------------
    def moe_dispatch(self, tokens: torch.Tensor, routing_map: torch.Tensor, probs: torch.Tensor, num_local_experts: int, topk: Optional[int] = None):
        """tokens [T, H] bf16, routing_map [T, E] bool, probs [T, E] → (handle, recv_tokens [M, H], recv_probs [M], tokens_per_local_expert [L] cpu)."""
        T, H = tokens.shape
        E, W, L = routing_map.shape[1], self.world, num_local_experts
        assert E == W * L, "experts must be evenly divided over the expert-parallel group"
        dev = tokens.device
        rmT = routing_map.bool().T.contiguous()                                    # [E, T] expert-major
        counts = rmT.sum(1).to(torch.int32)
        src_tok = torch.arange(T, device=dev).unsqueeze(0).expand(E, -1).masked_select(rmT)
        pair_exp = torch.arange(E, device=dev).unsqueeze(1).expand(-1, T).masked_select(rmT)
        all_counts = self.all_gather(counts.view(1, E)).clone().long()             # [W, E]
        tot = all_counts.sum(0).view(W, L)
        expert_base = (tot.cumsum(1) - tot).view(-1)                               # start row of each expert on its owner
        before_me = all_counts[: self.rank].sum(0)
        start = counts.long().cumsum(0) - counts.long()
        within = torch.arange(src_tok.numel(), device=dev) - start[pair_exp]
        dst_rank = (pair_exp // L).to(torch.int32)
        dst_slot = expert_base[pair_exp] + before_me[pair_exp] + within
        tpe = tot[self.rank].cpu()                                                 # the one host sync: sizes for the grouped GEMM
        m_recv = int(tpe.sum())
        if topk is None:
            topk = int(routing_map.sum(1).max())
        # per-token table of the same pairs (combine / dispatch-backward read through it)
        order = torch.argsort(src_tok, stable=True)
        st = src_tok[order]
        per_tok = torch.bincount(src_tok, minlength=T)
        tok_start = per_tok.cumsum(0) - per_tok
        j = torch.arange(st.numel(), device=dev) - tok_start[st]
        tok_rank = torch.zeros(T, topk, dtype=torch.int32, device=dev)
        tok_slot = torch.full((T, topk), -1, dtype=torch.int64, device=dev)
        tok_rank[st, j] = dst_rank[order]
        tok_slot[st, j] = dst_slot[order]
        h = DispatchHandle(T, topk, src_tok, dst_rank, dst_slot, tok_rank, tok_slot, m_recv)
        recv = _DispatchFn.apply(tokens, self, h)
        pair_probs = probs.T.contiguous().masked_select(rmT)
        recv_probs = _PairPushFn.apply(pair_probs, self, h)
        return h, recv, recv_probs, tpe

    def moe_combine(self, hidden: torch.Tensor, handle: DispatchHandle) -> torch.Tensor:
        """hidden [M, H] (expert outputs in receive order) → [T, H]: each token sums its k expert outputs."""
        return _CombineFn.apply(hidden, self, handle)


class _DispatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, be, h: DispatchHandle):
        ctx.be, ctx.h = be, h
        return be._moe_push(tokens, h.src_tok, h.dst_rank, h.dst_slot, h.m_recv)

    @staticmethod
    def backward(ctx, g):
        be, h = ctx.be, ctx.h
        return be._moe_pull(g, h.tok_rank, h.tok_slot, None, raw=False), None, None


class _CombineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, be, h: DispatchHandle):
        ctx.be, ctx.h = be, h
        return be._moe_pull(hidden, h.tok_rank, h.tok_slot, None, raw=False)

    @staticmethod
    def backward(ctx, g):
        be, h = ctx.be, ctx.h
        return be._moe_push(g.contiguous(), h.src_tok, h.dst_rank, h.dst_slot, h.m_recv), None, None


class _PairPushFn(torch.autograd.Function):
    """Per-pair scalars (routing probabilities) travel with the tokens: fp32 padded to one 16-byte vector per row."""

    @staticmethod
    def forward(ctx, pair_vals, be, h: DispatchHandle):
        ctx.be, ctx.h, ctx.dtype = be, h, pair_vals.dtype
        x = torch.zeros(pair_vals.numel(), 4, dtype=torch.float32, device=pair_vals.device)
        x[:, 0] = pair_vals.float()
        ar = torch.arange(pair_vals.numel(), device=pair_vals.device)
        return be._moe_push(x, ar, h.dst_rank, h.dst_slot, h.m_recv)[:, 0].to(pair_vals.dtype)

    @staticmethod
    def backward(ctx, g):
        be, h = ctx.be, ctx.h
        y = torch.zeros(g.numel(), 4, dtype=torch.float32, device=g.device)
        y[:, 0] = g.float()
        out = be._moe_pull(y, h.dst_rank.contiguous(), h.dst_slot.contiguous(), None, raw=True)
        return out[:, 0].to(ctx.dtype), None, None
