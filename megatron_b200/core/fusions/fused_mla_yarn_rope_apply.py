"""Rotary embedding for multi-latent attention, fused with the surrounding layout work (reference ``fusions/fused_mla_yarn_rope_apply.py``: Triton kernels
``fused_apply_mla_rope_for_q`` / ``fused_apply_mla_rope_for_kv``).

* query: the trailing ``emb`` channels of every head of ``q [s, b, n, nope + emb]`` are rotated IN PLACE (no split / rotate / cat round trip); DeepSeek
  checkpoints store the rotary pairs adjacent ("interleaved"), the output is half-split.
* key / value: ``kv [s, b, n, kd + vd]`` and the ONE rotary key ``k_pe [s, b, 1, emb]`` shared by the heads become ``key [s, b, n, kd + emb]`` and
  ``value [s, b, n, vd]`` in one pass (the rotation of ``k_pe`` included unless it is already rotated); the backward sums the rotary gradient over heads.

Angles are passed as angles (``[positions, emb]`` fp32; YaRN's ``mscale`` multiplies cos and sin).  CUDA: ``ops/csrc/routing_kernels.cu``; CPU: PyTorch."""
from typing import Optional

import torch

from ... import ops


def _kernels_ok(t: torch.Tensor, emb: int) -> bool:
    return t.is_cuda and ops.has_ext() and hasattr(ops.ext(), "mla_rope_inplace") and emb % 2 == 0 and emb <= 64 and t.dtype in (torch.bfloat16, torch.float16, torch.float32)


def _angles2d(angles: torch.Tensor) -> torch.Tensor:
    return angles.reshape(angles.shape[0], angles.shape[-1]).float().contiguous()


def _rotate_ref(x_pe, ang, mscale, interleaved, inverse=False):
    """x_pe [rows.., emb] fp32; ang broadcastable [.., emb]."""
    half = x_pe.shape[-1] // 2
    cos, sin = torch.cos(ang) * mscale, torch.sin(ang) * mscale
    if not inverse:
        a, b = (x_pe[..., 0::2], x_pe[..., 1::2]) if interleaved else (x_pe[..., :half], x_pe[..., half:])
        return torch.cat([a * cos[..., :half] - b * sin[..., :half], b * cos[..., half:] + a * sin[..., half:]], dim=-1)
    gl, gr = x_pe[..., :half], x_pe[..., half:]
    da, db = gl * cos[..., :half] + gr * sin[..., half:], gr * cos[..., half:] - gl * sin[..., :half]
    return torch.stack([da, db], dim=-1).flatten(-2) if interleaved else torch.cat([da, db], dim=-1)


class _MLARope(torch.autograd.Function):
    """One pass over q: untouched channels copied, rotary channels rotated.  ``inplace`` rewrites q itself (only legal when q is not a view of another
    custom Function's output — autograd forbids that combination)."""

    @staticmethod
    def forward(ctx, q, ang2d, positions, nope, emb, batch, mscale, interleaved, inplace):
        out = ops.ext().mla_rope_inplace(q, ang2d, positions, nope, emb, batch, mscale, interleaved, False, inplace)
        ops._count()
        if inplace:
            ctx.mark_dirty(q)
        ctx.save_for_backward(ang2d, *(() if positions is None else (positions,)))
        ctx.args = (nope, emb, batch, mscale, interleaved)
        return out

    @staticmethod
    def backward(ctx, g):
        ang2d, *rest = ctx.saved_tensors
        nope, emb, batch, mscale, interleaved = ctx.args
        gi = ops.ext().mla_rope_inplace(g.contiguous(), ang2d, rest[0] if rest else None, nope, emb, batch, mscale, interleaved, True, False)
        ops._count()
        return gi, None, None, None, None, None, None, None, None


def fused_apply_mla_rope_for_q(q: torch.Tensor, angles: torch.Tensor, nope_dim: int, emb_dim: int, mscale: float = 1.0, rotary_interleaved: bool = False,
                               position_ids: Optional[torch.Tensor] = None, inplace: bool = False) -> torch.Tensor:
    """``q [s, b, n, nope + emb]`` (or ``[t, n, nope + emb]`` with ``position_ids [t]``) → q with the rotary channels of every head rotated: one kernel, one
    pass (``inplace=True``: q itself is rewritten, the reference's ``fused_mla_rope_inplace``)."""
    if _kernels_ok(q, emb_dim) and q.is_contiguous():
        batch = q.shape[1] if q.dim() == 4 else 1
        pos = position_ids.long().contiguous() if position_ids is not None else None
        return _MLARope.apply(q, _angles2d(angles), pos, nope_dim, emb_dim, batch, float(mscale), bool(rotary_interleaved), bool(inplace))
    ang = _angles2d(angles)
    ang = ang[position_ids.long()] if position_ids is not None else ang[: q.shape[0]]
    ang = ang.view(ang.shape[0], *([1] * (q.dim() - 2)), emb_dim)
    rot = _rotate_ref(q[..., nope_dim:].float(), ang, mscale, rotary_interleaved).to(q.dtype)
    return torch.cat([q[..., :nope_dim], rot], dim=-1)


class _MLAKVSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kv, k_pe, ang2d, positions, kd, vd, emb, batch, mscale, interleaved):
        key, val = ops.ext().mla_kv_split(kv.contiguous(), k_pe.contiguous(), ang2d, positions, kd, vd, emb, batch, mscale, interleaved, False)
        ops._count()
        ctx.save_for_backward(*(t for t in (ang2d, positions) if t is not None))
        ctx.flags = (ang2d is not None, positions is not None)
        ctx.args = (kd, vd, emb, batch, mscale, interleaved)
        ctx.kpe_shape = k_pe.shape
        return key, val

    @staticmethod
    def backward(ctx, dkey, dval):
        saved = list(ctx.saved_tensors)
        ang2d = saved.pop(0) if ctx.flags[0] else None
        positions = saved.pop(0) if ctx.flags[1] else None
        kd, vd, emb, batch, mscale, interleaved = ctx.args
        dkv, dkpe = ops.ext().mla_kv_split(dkey.contiguous(), dval.contiguous(), ang2d, positions, kd, vd, emb, batch, mscale, interleaved, True)
        ops._count()
        return dkv, dkpe.view(ctx.kpe_shape), None, None, None, None, None, None, None, None


def fused_apply_mla_rope_for_kv(kv: torch.Tensor, k_pos_emb: torch.Tensor, angles: Optional[torch.Tensor], emb_dim: int, k_dim: int, v_dim: int, mscale: float = 1.0,
                                rotary_interleaved: bool = False, position_ids: Optional[torch.Tensor] = None):
    """``kv [s, b, n, k_dim + v_dim]`` + ``k_pos_emb [s, b, 1, emb]`` → ``(key [s, b, n, k_dim + emb], value [s, b, n, v_dim])``.  ``angles=None``: ``k_pos_emb``
    is already rotated (the split / broadcast / concatenate is still one kernel)."""
    if _kernels_ok(kv, emb_dim) and kv.dtype == k_pos_emb.dtype:
        batch = kv.shape[1] if kv.dim() == 4 else 1
        pos = position_ids.long().contiguous() if position_ids is not None else None
        return _MLAKVSplit.apply(kv, k_pos_emb, _angles2d(angles) if angles is not None else None, pos, k_dim, v_dim, emb_dim, batch, float(mscale), bool(rotary_interleaved))
    k_nope, v = torch.split(kv, [k_dim, v_dim], dim=-1)
    pe = k_pos_emb
    if angles is not None:
        ang = _angles2d(angles)
        ang = ang[position_ids.long()] if position_ids is not None else ang[: kv.shape[0]]
        ang = ang.view(ang.shape[0], *([1] * (kv.dim() - 2)), emb_dim)
        pe = _rotate_ref(pe.float(), ang, mscale, rotary_interleaved).to(kv.dtype)
    key = torch.cat([k_nope, pe.expand(*kv.shape[:-1], emb_dim)], dim=-1)
    return key, v.contiguous()
