"""Context parallelism for attention (reference: delegated to TransformerEngine).

Every CP rank holds two zig-zag chunks of the sequence (chunks ``r`` and ``2cp-1-r`` of ``2cp``), so the
causal work is balanced.  Two communication types:

* ``all_gather`` — all-gather K/V over the CP group, restore natural order, attend each local query chunk to
  its causal prefix.  Backward = reduce-scatter of dK/dV through the differentiable gather.
* ``p2p``        — ring attention: K/V blocks circulate with ``isend/irecv`` (next block prefetched while the
  current one is used); partial results are merged with the online-softmax (log-sum-exp) rule.  The block
  kernel returns (out, lse) and is differentiable in both, so autograd produces the ring backward.
* ``a2a``        — DeepSpeed-Ulysses: all-to-all swaps sequence sharding for head sharding around a plain
  full-sequence attention.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

from ..core import parallel_state as ps
from ..core.tensor_parallel.mappings import _AllToAll, _gather_along_first_dim, _reduce_scatter_along_first_dim


class _GatherSeq(torch.autograd.Function):
    """all-gather along dim 0 over ``group``; backward reduce-scatters."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _gather_along_first_dim(x.contiguous(), group)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_along_first_dim(g.contiguous(), ctx.group), None


def _zigzag_to_natural(full: torch.Tensor, cp: int) -> torch.Tensor:
    """[cp * 2 * c, ...] in rank order (r → chunks r, 2cp-1-r) → natural chunk order."""
    c = full.shape[0] // (2 * cp)
    chunks = full.view(cp, 2, c, *full.shape[1:])
    order = [None] * (2 * cp)
    for r in range(cp):
        order[r] = chunks[r, 0]
        order[2 * cp - 1 - r] = chunks[r, 1]
    return torch.cat(order, dim=0)


def attention_with_lse(q, k, v, scale: float, causal: bool, q_offset: int = 0, k_offset: int = 0):
    """Blockwise attention returning (out [sq,b,h,d] fp32, lse [b,h,sq] fp32); differentiable in both.

    ``q_offset``/``k_offset`` are the global positions of the first query/key (for the causal mask).
    GQA: k/v heads are repeated.  O(sq·sk) memory for the block — this is the portable ring building
    block (CPU tests + any GPU); the fused kernel replaces it on B200."""
    sq, b, hq, d = q.shape
    sk, hk = k.shape[0], k.shape[2]
    rep = hq // hk
    qf = q.permute(1, 2, 0, 3).float()
    kf = k.permute(1, 2, 0, 3).float()
    vf = v.permute(1, 2, 0, 3).float()
    if rep > 1:
        kf, vf = kf.repeat_interleave(rep, dim=1), vf.repeat_interleave(rep, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        qi = torch.arange(sq, device=q.device)[:, None] + q_offset
        ki = torch.arange(sk, device=q.device)[None, :] + k_offset
        s = s.masked_fill(ki > qi, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    lse_safe = torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)
    p = torch.exp(s - lse_safe.unsqueeze(-1))
    out = torch.matmul(p, vf).permute(2, 0, 1, 3)
    return out, lse


def merge_partials(outs: List[torch.Tensor], lses: List[torch.Tensor]) -> torch.Tensor:
    """Online-softmax merge: out = Σ_i exp(lse_i − lse) · out_i with lse = logsumexp_i lse_i."""
    L = torch.stack(lses)  # [n, b, h, sq]
    tot = torch.logsumexp(L, dim=0)
    w = torch.exp(L - tot.unsqueeze(0))  # [n, b, h, sq]
    w = torch.nan_to_num(w, nan=0.0)
    acc = 0
    for o, wi in zip(outs, w):
        acc = acc + o * wi.permute(2, 0, 1).unsqueeze(-1)
    return acc


# ---- ring attention on block kernels --------------------------------------------------------------------------------------------------------------------
# The ring is written against two block primitives so that the SAME schedule runs on the tcgen05 kernels (CUDA, bf16, head dim 128) and on a PyTorch
# formulation (CPU tests, other shapes):
#   block_fwd(q, k, v)                       -> (out, lse)          one (query chunk, key chunk) pair; chunks are aligned, so a pair is either fully visible
#   block_bwd(go, q, k, v, out_final, lse_final) -> (dq, dk, dv)    (non-causal), the diagonal (causal, sq == sk) or invisible (skipped)
# The backward of a pair uses the FINAL (merged) output and log-sum-exp of the query rows: P = exp(S - lse_final), dS = P o (dP - rowsum(dO o O_final)) — exactly
# what our native backward kernels take as inputs, so no partial results are kept from the forward.


def _native_block_ok(q, k) -> bool:
    from .. import ops

    return (q.is_cuda and ops.has_ext() and hasattr(ops.ext(), "flash_attn_bwd") and q.dtype == torch.bfloat16 and q.shape[-1] == 128 and k.shape[-1] == 128
            and q.shape[0] >= 128 and k.shape[0] >= 128 and q.shape[2] % k.shape[2] == 0)


def block_fwd(q, k, v, scale: float, causal: bool):
    """→ (out [sq, b, h, d] in q's dtype, lse [b, h, sq] fp32)."""
    if _native_block_ok(q, k):
        from .. import ops

        o, lse = ops.ext().flash_attn_fwd(q, k, v, causal, scale, 1)
        ops._count()
        return o, lse
    with torch.no_grad():
        o, lse = attention_with_lse(q, k, v, scale, causal)
    return o.to(q.dtype), lse


def block_bwd(go, q, k, v, out_final, lse_final, scale: float, causal: bool):
    """→ (dq, dk, dv) of one pair given the merged output / log-sum-exp of its query rows."""
    if _native_block_ok(q, k):
        from .. import ops

        dq, dk, dv = ops.ext().flash_attn_bwd(go.contiguous(), q, k, v, out_final.contiguous(), lse_final.contiguous(), causal, scale, -1)
        ops._count(3)
        return dq, dk, dv
    sq, b, hq, d = q.shape
    sk, hk = k.shape[0], k.shape[2]
    rep = hq // hk
    qf, gof, of = (t.permute(1, 2, 0, 3).float() for t in (q, go, out_final))
    kf, vf = k.permute(1, 2, 0, 3).float(), v.permute(1, 2, 0, 3).float()
    if rep > 1:
        kf, vf = kf.repeat_interleave(rep, dim=1), vf.repeat_interleave(rep, dim=1)
    sc = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        sc = sc.masked_fill(torch.ones(sq, sk, dtype=torch.bool, device=q.device).triu(1), float("-inf"))
    pr = torch.exp(sc - lse_final.unsqueeze(-1))
    dv = torch.matmul(pr.transpose(-1, -2), gof)
    ds = pr * (torch.matmul(gof, vf.transpose(-1, -2)) - (gof * of).sum(-1, keepdim=True)) * scale
    dq = torch.matmul(ds, kf)
    dk = torch.matmul(ds.transpose(-1, -2), qf)
    if rep > 1:
        dk, dv = dk.view(b, hk, rep, sk, d).sum(2), dv.view(b, hk, rep, sk, v.shape[-1]).sum(2)
    return dq.permute(2, 0, 1, 3).to(q.dtype), dk.permute(2, 0, 1, 3).to(k.dtype), dv.permute(2, 0, 1, 3).to(v.dtype)


def _merge_into(acc_o, acc_l, o, l):
    """Online-softmax merge of one more partial (o in any dtype, accumulators fp32)."""
    if acc_o is None:
        return o.float(), l.clone()
    new_l = torch.logaddexp(acc_l, l)
    wa, wb = torch.exp(acc_l - new_l).permute(2, 0, 1).unsqueeze(-1), torch.exp(l - new_l).permute(2, 0, 1).unsqueeze(-1)
    return acc_o * wa + o.float() * wb, new_l


class _RingAttnFn(torch.autograd.Function):
    """Zig-zag ring attention: K/V blocks circulate (next hop in flight while the current block is used); in the backward dK/dV travel WITH their block and arrive
    home after a full turn.  ``shift(x, reverse)`` moves a tensor one hop; ``rank`` / ``cp`` place the chunks."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal, rank, cp, shift):
        c = q.shape[0] // 2
        my_offs = [rank * c, (2 * cp - 1 - rank) * c]
        acc = [[None, None], [None, None]]
        cur = torch.cat([k, v], dim=-1)
        dk_ = k.shape[-1]
        overlap = hasattr(shift, "start")
        for step in range(cp):
            src = (rank - step) % cp
            nxt = (shift.start(cur, False) if overlap else shift(cur, False)) if step < cp - 1 else None      # next block in flight while this one is used
            src_offs = [src * c, (2 * cp - 1 - src) * c]
            for qi, qoff in enumerate(my_offs):
                for ki, koff in enumerate(src_offs):
                    if causal and koff > qoff:
                        continue
                    o, l = block_fwd(q[qi * c:(qi + 1) * c], cur[ki * c:(ki + 1) * c, ..., :dk_], cur[ki * c:(ki + 1) * c, ..., dk_:], scale, causal and koff == qoff)
                    acc[qi] = list(_merge_into(acc[qi][0], acc[qi][1], o, l))
            cur = shift.finish(nxt) if (overlap and nxt is not None) else nxt
        out = torch.cat([acc[0][0], acc[1][0]], dim=0).to(q.dtype)
        lse = torch.cat([acc[0][1], acc[1][1]], dim=-1)                       # [b, h, 2c]
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.args = (scale, causal, rank, cp, shift)
        return out

    @staticmethod
    def backward(ctx, go):
        q, k, v, out, lse = ctx.saved_tensors
        scale, causal, rank, cp, shift = ctx.args
        c = q.shape[0] // 2
        my_offs = [rank * c, (2 * cp - 1 - rank) * c]
        go = go.contiguous()
        dq = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
        dk_ = k.shape[-1]
        cur = torch.cat([k, v], dim=-1)
        dcur = torch.zeros(cur.shape, dtype=torch.float32, device=q.device)     # gradient of the block currently held; travels with it
        overlap = hasattr(shift, "start")
        pending_d = None                                                        # the gradient hop of the previous step, still in flight
        for step in range(cp):
            src = (rank - step) % cp
            nxt = (shift.start(cur, False) if overlap else shift(cur, False)) if step < cp - 1 else None
            src_offs = [src * c, (2 * cp - 1 - src) * c]
            contrib = []
            for qi, qoff in enumerate(my_offs):
                qs = slice(qi * c, (qi + 1) * c)
                for ki, koff in enumerate(src_offs):
                    if causal and koff > qoff:
                        continue
                    ks = slice(ki * c, (ki + 1) * c)
                    g_q, g_k, g_v = block_bwd(go[qs], q[qs], cur[ks][..., :dk_], cur[ks][..., dk_:], out[qs], lse[..., qs].contiguous(), scale, causal and koff == qoff)
                    dq[qs] += g_q.float()
                    contrib.append((ks, g_k, g_v))
            if pending_d is not None:
                dcur = shift.finish(pending_d)                                  # arrives while this step's pairs were being computed
            for ks, g_k, g_v in contrib:
                dcur[ks, ..., :dk_] += g_k.float()
                dcur[ks, ..., dk_:] += g_v.float()
            # the gradient follows its block: one hop per step, and a last hop after the final step brings it back to the block's owner
            if overlap:
                pending_d = shift.start(dcur, False)
                cur = shift.finish(nxt) if nxt is not None else None
            else:
                dcur = shift(dcur, False)
                cur = nxt
        if pending_d is not None:
            dcur = shift.finish(pending_d)
        return dq.to(q.dtype), dcur[..., :dk_].to(k.dtype), dcur[..., dk_:].to(v.dtype), None, None, None, None, None


class _RingShift(torch.autograd.Function):
    """Send a tensor to the next CP rank and receive from the previous one (backward: the reverse)."""

    @staticmethod
    def forward(ctx, x, group, reverse):
        ctx.group, ctx.reverse = group, reverse
        return _RingShift._shift(x, group, reverse)

    @staticmethod
    def _shift(x, group, reverse):
        ranks = dist.get_process_group_ranks(group)
        me = dist.get_rank(group)
        n = len(ranks)
        dst = ranks[(me - 1) % n] if reverse else ranks[(me + 1) % n]
        src = ranks[(me + 1) % n] if reverse else ranks[(me - 1) % n]
        x = x.contiguous()
        out = torch.empty_like(x)
        if me % 2 == 0:
            reqs = [dist.isend(x, dst, group=group), dist.irecv(out, src, group=group)]
        else:
            reqs = [dist.irecv(out, src, group=group), dist.isend(x, dst, group=group)]
        for r in reqs:
            r.wait()
        return out

    @staticmethod
    def backward(ctx, g):
        return _RingShift._shift(g, ctx.group, not ctx.reverse), None, None

    @staticmethod
    def start(x, group, reverse=False):
        """Post the send / receive of one hop and return ``(receive buffer, requests)`` without waiting: the transfer runs on the communicator's stream while the
        caller computes on the block it already holds."""
        ranks = dist.get_process_group_ranks(group)
        me = dist.get_rank(group)
        n = len(ranks)
        dst = ranks[(me - 1) % n] if reverse else ranks[(me + 1) % n]
        src = ranks[(me + 1) % n] if reverse else ranks[(me - 1) % n]
        x = x.contiguous()
        out = torch.empty_like(x)
        ops_ = [dist.P2POp(dist.isend, x, dst, group), dist.P2POp(dist.irecv, out, src, group)]
        return out, dist.batch_isend_irecv(ops_), x          # x is kept alive until the send has completed

    @staticmethod
    def finish(handle):
        out, reqs, _keep = handle
        for r in reqs:
            r.wait()
        return out


class _AsyncRing:
    """``shift`` object for ``_RingAttnFn``: ``start`` / ``finish`` overlap each hop with the block computation; calling it does a blocking hop."""

    def __init__(self, group):
        self.group = group

    def __call__(self, x, reverse=False):
        return _RingShift.finish(_RingShift.start(x, self.group, reverse))

    def start(self, x, reverse=False):
        return _RingShift.start(x, self.group, reverse)

    finish = staticmethod(_RingShift.finish)


class RingAttention(torch.nn.Module):
    """Drop-in core attention for ``context_parallel_size > 1`` (``DotProductAttention`` delegates here)."""

    def __init__(self, config, cp_comm_type: str = "p2p", pg_collection=None):
        super().__init__()
        self.config = config
        self.kind = cp_comm_type if isinstance(cp_comm_type, str) else "p2p"
        self.group = pg_collection.cp if (pg_collection is not None and getattr(pg_collection, "cp", None) is not None) else ps.get_context_parallel_group()
        self.cp = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)

    def _chunk_offsets(self, s_local: int) -> Tuple[int, List[int]]:
        c = s_local // 2
        return c, [self.rank * c, (2 * self.cp - 1 - self.rank) * c]

    def forward(self, q, k, v, causal: bool, scale: float):
        if self.kind in ("all_gather",):
            return self._all_gather(q, k, v, causal, scale)
        if self.kind in ("a2a",):
            return self._ulysses(q, k, v, causal, scale)
        if self.kind in ("a2a+p2p",):
            return self._hierarchical(q, k, v, causal, scale)
        return self._ring(q, k, v, causal, scale)

    # ---- hierarchical: all-to-all over heads inside the inner group, ring over the outer group --------------------
    def _hierarchical(self, q, k, v, causal, scale):
        """``cp = a · p`` with ``hierarchical_context_parallel_sizes = [a, p]`` (reference ``cp_comm_type="a2a+p2p"``): the ``a`` ranks of an inner
        group (one NVLink island) trade sequence for heads (Ulysses), so each holds ``h/a`` heads of the ``2a`` zig-zag chunks of its island; the
        key/value blocks then travel around the ring of ``p`` islands.  Only ``p - 1`` ring steps with ``1/a`` of the heads each, instead of ``cp - 1``."""
        groups = ps.get_hierarchical_context_parallel_groups()
        g_in, g_out = groups[0], groups[1]
        a, p_sz = dist.get_world_size(g_in), dist.get_world_size(g_out)
        r_out = dist.get_rank(g_out)
        s, b, h, d = q.shape
        assert h % a == 0 and k.shape[2] % a == 0, "a2a+p2p needs query and key/value heads divisible by the inner context-parallel size"
        c = s // 2

        def seq_to_head(t):
            hh = t.shape[2]
            x = t.view(s, b, a, hh // a, t.shape[3]).permute(2, 0, 1, 3, 4).contiguous().view(a * s, b, hh // a, t.shape[3])
            return _AllToAll.apply(g_in, x, None, None)          # blocks ordered by source inner rank, each [2c, ...] = (low chunk, high chunk)

        def offsets(outer):                                     # global token offsets of the 2a chunks held by island ``outer`` (inner rank fastest)
            offs = []
            for i in range(a):
                r = outer * a + i
                offs += [r * c, (2 * self.cp - 1 - r) * c]
            return offs

        qf, kf, vf = seq_to_head(q), seq_to_head(k), seq_to_head(v)
        dk = kf.shape[-1]
        cur = torch.cat([kf, vf], dim=-1)
        q_offs = offsets(r_out)
        outs = [[] for _ in range(2 * a)]
        lses = [[] for _ in range(2 * a)]
        for step in range(p_sz):
            src = (r_out - step) % p_sz
            nxt = _RingShift.apply(cur, g_out, False) if step < p_sz - 1 else None
            kk, vv = cur[..., :dk], cur[..., dk:]
            k_offs = offsets(src)
            for qi, qoff in enumerate(q_offs):
                qc = qf[qi * c : (qi + 1) * c]
                for ki, koff in enumerate(k_offs):
                    if causal and koff > qoff + c - 1:
                        continue
                    o, l = attention_with_lse(qc, kk[ki * c : (ki + 1) * c], vv[ki * c : (ki + 1) * c], scale, causal, q_offset=qoff, k_offset=koff)
                    outs[qi].append(o), lses[qi].append(l)
            cur = nxt
        merged = torch.cat([merge_partials(outs[i], lses[i]) for i in range(2 * a)], dim=0).to(q.dtype)     # [a * s, b, h/a, dv]
        y = _AllToAll.apply(g_in, merged.contiguous(), None, None)
        return y.view(a, s, b, h // a, y.shape[-1]).permute(1, 2, 0, 3, 4).reshape(s, b, h, y.shape[-1])

    # ---- all-gather KV ------------------------------------------------------------------------------------
    def _all_gather(self, q, k, v, causal, scale):
        c, offs = self._chunk_offsets(q.shape[0])
        kf = _zigzag_to_natural(_GatherSeq.apply(k, self.group), self.cp)
        vf = _zigzag_to_natural(_GatherSeq.apply(v, self.group), self.cp)
        outs = []
        for i, off in enumerate(offs):
            qc = q[i * c : (i + 1) * c]
            end = off + c if causal else kf.shape[0]
            o, _ = attention_with_lse(qc, kf[:end], vf[:end], scale, causal, q_offset=off, k_offset=0)
            outs.append(o)
        return torch.cat(outs, dim=0).to(q.dtype)

    # ---- ring (p2p) -------------------------------------------------------------------------------------------
    def _ring(self, q, k, v, causal, scale):
        return _RingAttnFn.apply(q, k.contiguous(), v.contiguous(), scale, causal, self.rank, self.cp, _AsyncRing(self.group))

    # ---- Ulysses (a2a) --------------------------------------------------------------------------------------------
    def _ulysses(self, q, k, v, causal, scale):
        """[s/cp, b, h, d] → all-to-all → [s, b, h/cp, d] → attention → all-to-all back."""
        from .. import ops

        cp = self.cp

        def seq_to_head(t):
            s, b, h, d = t.shape
            assert h % cp == 0, "Ulysses needs heads divisible by the context-parallel size"
            x = t.view(s, b, cp, h // cp, d).permute(2, 0, 1, 3, 4).contiguous().view(cp * s, b, h // cp, d)
            y = _AllToAll.apply(self.group, x, None, None)
            return _zigzag_to_natural(y, cp)

        def head_to_seq(t, s_local):
            # inverse of the above: natural → zig-zag rank order, then a2a
            c = s_local // 2
            chunks = t.view(2 * cp, c, *t.shape[1:])
            order = []
            for r in range(cp):
                order += [chunks[r], chunks[2 * cp - 1 - r]]
            x = torch.cat(order, dim=0).contiguous()
            y = _AllToAll.apply(self.group, x, None, None)
            s, b, hh, d = s_local, y.shape[1], y.shape[2], y.shape[3]
            return y.view(cp, s, b, hh, d).permute(1, 2, 0, 3, 4).reshape(s, b, cp * hh, d)

        qf, kf, vf = seq_to_head(q), seq_to_head(k), seq_to_head(v)
        o = ops.flash_attention(qf, kf, vf, causal=causal, scale=scale)
        return head_to_seq(o, q.shape[0])


# ---- sequence ⇄ channel re-sharding for recurrent layers (Mamba context parallel; reference ``ssm/mamba_context_parallel.py``) -----------------------------------
def seq_to_channel(t: torch.Tensor, group, cp: int) -> torch.Tensor:
    """``[l/cp (zig-zag shard), b, C]`` → ``[l (natural order), b, C/cp]``: this rank ends up with the WHOLE sequence for its block of channels, which is
    what a causal conv / selective scan needs (they are sequential along l but independent across channels)."""
    l, b, C = t.shape
    assert C % cp == 0
    x = t.view(l, b, cp, C // cp).permute(2, 0, 1, 3).contiguous().view(cp * l, b, C // cp)
    return _zigzag_to_natural(_AllToAll.apply(group, x, None, None), cp)


def channel_to_seq(t: torch.Tensor, group, cp: int) -> torch.Tensor:
    """Inverse of ``seq_to_channel``: ``[l, b, C/cp]`` → ``[l/cp (zig-zag shard), b, C]``."""
    L, b, Cc = t.shape
    c = L // (2 * cp)
    chunks = t.view(2 * cp, c, b, Cc)
    order = []
    for r in range(cp):
        order += [chunks[r], chunks[2 * cp - 1 - r]]
    y = _AllToAll.apply(group, torch.cat(order, dim=0).contiguous(), None, None)          # [cp (source = channel block), l/cp, b, C/cp]
    return y.view(cp, 2 * c, b, Cc).permute(1, 2, 0, 3).reshape(2 * c, b, cp * Cc)
