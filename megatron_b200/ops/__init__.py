"""Operator library: hand-written sm_100a kernels with PyTorch references.

Every public function dispatches on the device of its input:

* CUDA tensor → the in-tree extension ``megatron_b200/ops/_C*.so`` (built by
  ``megatron_b200.ops.build``; sources in ``ops/csrc``).  A CUDA tensor with no
  extension available raises — there is no silent eager fallback on a GPU box
  (set ``MEGATRON_B200_ALLOW_FALLBACK=1`` to override for debugging).
* CPU tensor → ``ops.reference`` (plain PyTorch).

Replaces what the reference reaches through TransformerEngine / Apex /
``torch.compile``.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import reference as ref

_EXT = None
_EXT_ERR = None


def _load_ext():
    global _EXT, _EXT_ERR
    if _EXT is not None or _EXT_ERR is not None:
        return _EXT
    try:
        from . import _C  # type: ignore

        _EXT = _C
    except Exception as e:  # pragma: no cover - depends on build
        _EXT_ERR = e
    return _EXT


def ext():
    """The native extension module (raises if it is not built/loadable)."""
    e = _load_ext()
    if e is None:
        raise RuntimeError(
            f"megatron_b200 native extension is not available ({_EXT_ERR!r}); run `python -m megatron_b200.ops.build`"
        )
    return e


def has_ext() -> bool:
    return _load_ext() is not None


def _use_cuda(t: torch.Tensor) -> bool:
    if not t.is_cuda:
        return False
    if _load_ext() is None:
        if os.environ.get("MEGATRON_B200_ALLOW_FALLBACK") == "1":
            return False
        raise RuntimeError(
            f"CUDA tensor passed to megatron_b200.ops but the native extension failed to load: {_EXT_ERR!r}"
        )
    return True


# launch counter: bench.py reports how many of *our* kernels ran in the timed region
_LAUNCHES = 0


def _count(n=1):
    global _LAUNCHES
    _LAUNCHES += n


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count():
    global _LAUNCHES
    _LAUNCHES = 0


# =============================================================================
# GEMM (see ops/gemm.py: tcgen05 variants + cuBLAS, autotuned per problem)
# =============================================================================
from .gemm import gemm_nn, gemm_nt, gemm_tn, get_gemm_backend, set_gemm_backend  # noqa: E402


# =============================================================================
# Norms
# =============================================================================


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps, zero_centered):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if _use_cuda(x):
            x2 = x2.contiguous()
            y, rstd = ext().rmsnorm_fwd(x2, w, eps, zero_centered)
            _count()
        else:
            y, rstd = ref.rms_norm_fwd(x2, w, eps, zero_centered)
        ctx.save_for_backward(x2, w, rstd)
        ctx.zero_centered, ctx.shape = zero_centered, shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w, rstd = ctx.saved_tensors
        g2 = gy.reshape(x2.shape)
        if _use_cuda(g2):
            gx, gw = ext().rmsnorm_bwd(g2.contiguous(), x2, w, rstd, ctx.zero_centered)
            _count(2)
        else:
            gx, gw = ref.rms_norm_bwd(g2, x2, w, rstd, ctx.zero_centered)
        return gx.view(ctx.shape), gw, None, None


def rms_norm(x, weight, eps: float = 1e-5, zero_centered_gamma: bool = False):
    return _RMSNormFn.apply(x, weight, eps, zero_centered_gamma)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, zero_centered):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if _use_cuda(x):
            x2 = x2.contiguous()
            y, mu, rstd = ext().layernorm_fwd(x2, w, b, eps, zero_centered)
            _count()
        else:
            y, mu, rstd = ref.layer_norm_fwd(x2, w, b, eps, zero_centered)
        ctx.save_for_backward(x2, w, mu, rstd)
        ctx.zero_centered, ctx.shape, ctx.has_bias = zero_centered, shape, b is not None
        return y.view(shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w, mu, rstd = ctx.saved_tensors
        g2 = gy.reshape(x2.shape)
        if _use_cuda(g2):
            gx, gw, gb = ext().layernorm_bwd(g2.contiguous(), x2, w, mu, rstd, ctx.zero_centered)
            _count(2)
            if not ctx.has_bias:
                gb = None
        else:
            gx, gw, gb = ref.layer_norm_bwd(g2, x2, w, mu, rstd, ctx.zero_centered, ctx.has_bias)
        return gx.view(ctx.shape), gw, gb, None, None


def layer_norm(x, weight, bias, eps: float = 1e-5, zero_centered_gamma: bool = False):
    return _LayerNormFn.apply(x, weight, bias, eps, zero_centered_gamma)


# =============================================================================
# Gated activations
# =============================================================================


class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, bias, probs):
        ctx.save_for_backward(y, bias, probs)
        if _use_cuda(y):
            y2 = y.reshape(-1, y.shape[-1]).contiguous()
            out = ext().swiglu_fwd(y2, bias, probs.reshape(-1) if probs is not None else None)
            _count()
            return out.view(*y.shape[:-1], y.shape[-1] // 2)
        return ref.swiglu_fwd(y, bias, probs)

    @staticmethod
    def backward(ctx, g):
        y, bias, probs = ctx.saved_tensors
        if _use_cuda(g):
            y2 = y.reshape(-1, y.shape[-1]).contiguous()
            g2 = g.reshape(-1, g.shape[-1]).contiguous()
            dy, dprobs = ext().swiglu_bwd(g2, y2, bias, probs.reshape(-1) if probs is not None else None)
            _count()
            dy = dy.view(y.shape)
            if dprobs is not None:
                dprobs = dprobs.view(probs.shape)
        else:
            dy, dprobs = ref.swiglu_bwd(g, y, bias, probs)
        gb = dy.reshape(-1, dy.shape[-1]).sum(0).to(bias.dtype) if bias is not None else None
        return dy, gb, dprobs


def swiglu(y, bias=None, probs=None):
    """``silu(a) * b`` with ``a, b = (y + bias).chunk(2, -1)``; optional per-token ``probs`` weight (MoE)."""
    return _SwiGLUFn.apply(y, bias, probs)


def bias_swiglu(y, bias):
    return _SwiGLUFn.apply(y, bias, None)


def geglu(y, bias=None):
    yb = y if bias is None else y + bias
    a, b = yb.chunk(2, dim=-1)
    return torch.nn.functional.gelu(a, approximate="tanh") * b


def bias_gelu(y, bias=None):
    yb = y if bias is None else y + bias
    return torch.nn.functional.gelu(yb, approximate="tanh")


# =============================================================================
# RoPE
# =============================================================================


class _RoPEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, freqs, interleaved, mscale):
        ctx.save_for_backward(freqs)
        ctx.interleaved, ctx.mscale = interleaved, mscale
        if _use_cuda(t) and not interleaved and t.dim() == 4:
            out = ext().rope_fwd(t.contiguous(), freqs.reshape(freqs.shape[0], -1).float().contiguous(), float(mscale), False)
            _count()
            return out
        return ref.rope_fwd(t, freqs, interleaved, mscale, conj=False)

    @staticmethod
    def backward(ctx, g):
        (freqs,) = ctx.saved_tensors
        if _use_cuda(g) and not ctx.interleaved and g.dim() == 4:
            out = ext().rope_fwd(g.contiguous(), freqs.reshape(freqs.shape[0], -1).float().contiguous(), float(ctx.mscale), True)
            _count()
            return out, None, None, None
        return ref.rope_fwd(g, freqs, ctx.interleaved, ctx.mscale, conj=True), None, None, None


def apply_rope(t, freqs, interleaved: bool = False, mscale: float = 1.0):
    """Rotary embedding on ``t[s, b, h, d]`` with angles ``freqs[s, 1, 1, d_rot]``."""
    return _RoPEFn.apply(t, freqs, interleaved, mscale)


# =============================================================================
# bias + dropout + residual add
# =============================================================================


class _BiasDropoutAddFn(torch.autograd.Function):
    """``residual + dropout(x + bias)`` in one pass (``csrc/misc_kernels.cu``); the keep mask is a function of the generator's (seed, offset) at call time and is
    regenerated in the backward — nothing but two integers is saved."""

    @staticmethod
    def forward(ctx, x, bias, residual, prob):
        y, seed, offset = ext().bias_dropout_add_fwd(x, bias, residual, prob)
        _count()
        ctx.prob, ctx.seed, ctx.offset, ctx.has_bias = prob, seed, offset, bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gx = ext().bias_dropout_add_bwd(g, ctx.prob, ctx.seed, ctx.offset) if ctx.prob > 0 else g
        if ctx.prob > 0:
            _count()
        gb = gx.reshape(-1, gx.shape[-1]).sum(0) if ctx.has_bias and ctx.needs_input_grad[1] else None
        return gx, gb, g, None


def bias_dropout_add(x, bias, residual, prob: float, training: bool):
    p = float(prob) if training else 0.0
    if (_use_cuda(x) and (p > 0.0 or bias is not None) and hasattr(ext(), "bias_dropout_add_fwd") and x.dtype in (torch.bfloat16, torch.float16, torch.float32)
            and x.dtype == residual.dtype and x.shape == residual.shape and x.is_contiguous() and residual.is_contiguous() and x.shape[-1] % 8 == 0
            and (bias is None or (bias.dtype == x.dtype and bias.numel() == x.shape[-1] and bias.is_contiguous()))
            and x.data_ptr() % 16 == 0 and residual.data_ptr() % 16 == 0 and (bias is None or bias.data_ptr() % 16 == 0)
            and not (p > 0.0 and torch.cuda.is_current_stream_capturing())):
        return _BiasDropoutAddFn.apply(x, bias, residual, p)
    if bias is not None:
        x = x + bias
    if prob > 0.0 and training:
        x = torch.nn.functional.dropout(x, p=prob, training=True)
    return residual + x


# =============================================================================
# Attention
# =============================================================================


class _FlashAttnFn(torch.autograd.Function):
    """sm_100a flash attention (``csrc/flash_attn_sm100.cu``); q/k/v layout [s, b, h, d]."""

    @staticmethod
    def forward(ctx, q, k, v, causal, scale, row_lo=None, col_hi=None):
        o, lse = ext().flash_attn_fwd(q, k, v, causal, scale, _FA_VARIANT, row_lo)
        _count()
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal, ctx.scale, ctx.band = causal, scale, (row_lo, col_hi)
        return o

    @staticmethod
    def backward(ctx, go):
        q, k, v, o, lse = ctx.saved_tensors
        # backward = cuDNN library kernel fed with OUR forward's (out, log-sum-exp); a native tcgen05 backward is the next step
        go = go.contiguous()
        if _ATTN_BWD_IMPL == "native" and hasattr(ext(), "flash_attn_bwd") and q.shape[-1] == 128:
            # our tcgen05 backward (csrc/flash_attn_bwd_sm100.cu): dK/dV accumulate in TMEM while dS tiles stream to a bf16 scratch by TMA; dQ = dS K in a second tcgen05 kernel
            try:
                dq, dk, dv = ext().flash_attn_bwd(go, q, k, v, o, lse, ctx.causal, ctx.scale, _ATTN_BWD_SPLIT_HEADS, 0, ctx.band[0], ctx.band[1])
            except RuntimeError as e:
                raise RuntimeError(f"{e}; go {tuple(go.shape)} {go.stride()} q {q.stride()} k {k.stride()} v {v.stride()} o {o.stride()} ptrs {[t.data_ptr() % 16 for t in (go, q, k, v, o)]}") from e
            _count(4 if _ATTN_BWD_SPLIT_HEADS != 0 else 3)
            return dq, dk, dv, None, None, None, None
        assert ctx.band[0] is None, "band masks (sliding window / packed sequences) need the native backward (head dim 128)"
        lse_lib = lse.unsqueeze(-1) if _cudnn_lse_ndim() == 4 else lse
        dq, dk, dv = ext().attn_bwd_cudnn(go, q, k, v, o, lse_lib, ctx.causal, ctx.scale)
        return dq, dk, dv, None, None, None, None


_CUDNN_LSE_NDIM = None


def _cudnn_lse_ndim() -> int:
    """Rank of the log-sum-exp tensor this torch build's cuDNN SDPA produces ([b,h,s] or [b,h,s,1])."""
    global _CUDNN_LSE_NDIM
    if _CUDNN_LSE_NDIM is None:
        t = torch.zeros(1, 1, 128, 64, device="cuda", dtype=torch.bfloat16)
        _CUDNN_LSE_NDIM = torch.ops.aten._scaled_dot_product_cudnn_attention(t, t, t, None, True, 0.0, True)[1].dim()
    return _CUDNN_LSE_NDIM


_ATTN_IMPL = os.environ.get("MEGATRON_B200_ATTN", "auto")  # auto | native | library
_ATTN_BWD_IMPL = os.environ.get("MEGATRON_B200_ATTN_BWD", "native")  # native: our two-kernel tcgen05 backward (csrc/flash_attn_bwd_sm100.cu) | library: cuDNN on our (out, LSE)
_ATTN_BWD_SPLIT_HEADS = int(os.environ.get("MEGATRON_B200_ATTN_BWD_SPLIT_HEADS", "-1"))  # -1: one CTA per (key block, QUERY head) whenever the model has GQA (measured faster at every head count: 4x the CTAs, better tail balance); 0: one CTA per KV head
_FA_VARIANT = int(os.environ.get("MEGATRON_B200_FA_VARIANT", "1"))  # 1 (default, measured 836 vs 786 TF): P kept in tensor memory (TS MMA); 0: P through shared memory


# what "auto" means on this build: the faster MEASURED forward at the Llama-3 8B shape (profiles/r1_attention.md)
_ATTN_AUTO_RESOLVES_TO = "native"   # our tcgen05 kernels at every TP size (few-head grids use mirrored tile pairing / split-heads); "library" = cuDNN, for comparisons


def _resolved_attn_impl() -> str:
    return _ATTN_AUTO_RESOLVES_TO if _ATTN_IMPL == "auto" else _ATTN_IMPL


def attention_impl_for(sq: int, b: int, hq_local: int) -> str:
    """What ``flash_attention`` will run for a bf16 [sq, b, hq_local, 128] causal problem on this build (for logs / bench reports)."""
    impl = _resolved_attn_impl()
    if impl == "native":
        grid = ((sq + 255) // 256) * hq_local * b
        fwd = "tcgen05 forward (ours, P in TMEM" + (", mirrored tile pairing" if grid <= 148 * 3 // 2 else "") + ")"
        return fwd + (" + tcgen05 backward (ours: dK/dV kernel + dQ=dS.K kernel)" if _ATTN_BWD_IMPL == "native" else " + cuDNN backward on our (out, LSE)")
    return "cuDNN SDPA (library) forward + backward"


def set_attention_impl(impl: str) -> None:
    global _ATTN_IMPL
    assert impl in ("auto", "native", "library")
    _ATTN_IMPL = impl


_BAND_CACHE: dict = {}


def attention_band(sq: int, sk: int, window=None, cu_seqlens=None, device="cpu"):
    """The monotone band of a causal mask with a sliding window and / or packed sequences, as the two int32 arrays our kernels take:
    ``row_lo[q]`` = first key query q sees, ``col_hi[k]`` = one past the last query that sees key k (the other edge of both is the causal diagonal
    ``k <= q + sk - sq``).  ``window = (left, right)`` in the reference's convention (q sees keys ``>= q + off - left``; ``left < 0`` = unbounded);
    ``cu_seqlens`` = cumulative lengths of sequences packed along the token dim (self-attention: the same boundaries for queries and keys).
    Returns ``None`` when there is nothing to restrict."""
    left = window[0] if window is not None and window[0] is not None and window[0] >= 0 else None
    if left is None and cu_seqlens is None:
        return None
    key = (sq, sk, left, str(device), None if cu_seqlens is None else (cu_seqlens.data_ptr(), cu_seqlens._version, cu_seqlens.numel()))
    hit = _BAND_CACHE.get(key)
    if hit is not None:
        return hit[0], hit[1]
    off = sk - sq
    qpos, kpos = torch.arange(sq, device=device), torch.arange(sk, device=device)
    row_lo, col_hi = torch.zeros(sq, dtype=torch.long, device=device), torch.full((sk,), sq, dtype=torch.long, device=device)
    if left is not None:
        row_lo = torch.maximum(row_lo, qpos + off - left)
        col_hi = torch.minimum(col_hi, kpos - off + left + 1)
    if cu_seqlens is not None:
        assert sq == sk, "packed sequences: self-attention only (same boundaries for queries and keys)"
        cu = cu_seqlens.to(device=device, dtype=torch.long)
        sid = torch.bucketize(qpos, cu[1:], right=True).clamp_(max=cu.numel() - 2)     # trailing pad tokens join the last sequence
        row_lo = torch.maximum(row_lo, cu[sid])
        col_hi = torch.minimum(col_hi, torch.where(sid == cu.numel() - 2, torch.full_like(sid, sq), cu[sid + 1]))
    band = (row_lo.clamp_(min=0).to(torch.int32).contiguous(), col_hi.clamp_(min=0, max=sq).to(torch.int32).contiguous())
    if len(_BAND_CACHE) > 64:
        _BAND_CACHE.clear()
    _BAND_CACHE[key] = band + (cu_seqlens,)          # keep the key tensor alive: its data_ptr is part of the key
    return band


def _native_attention_ok(q, k, v, causal, window, cu_seqlens=None) -> bool:
    if _resolved_attn_impl() == "library" or not hasattr(ext(), "flash_attn_fwd"):
        return False
    banded = cu_seqlens is not None or (window is not None and window[0] is not None and window[0] >= 0)
    if banded:                                                  # band masks: causal only, native backward only (head dim 128), packed = one batch row
        if not causal or q.shape[-1] != 128 or _ATTN_BWD_IMPL != "native" or not hasattr(ext(), "flash_attn_bwd"):
            return False
        if (window is not None and len(window) > 1 and window[1] not in (0, -1, None)) or (cu_seqlens is not None and (q.shape[1] != 1 or q.shape[0] != k.shape[0])):
            return False
    if q.dtype != torch.bfloat16 or q.shape[-1] not in (64, 128) or k.shape[-1] != q.shape[-1] or v.shape[-1] != q.shape[-1]:
        return False
    if (causal and k.shape[0] < q.shape[0]) or q.shape[0] < 128:   # decode-sized queries stay on the library path (untested regime for the 2x128-row tiling)
        return False
    strides_ok = all(t.stride(-1) == 1 and all(s % 8 == 0 for s in t.stride()[:-1]) and t.data_ptr() % 16 == 0 for t in (q, k, v))
    return strides_ok


def flash_attention(q, k, v, causal: bool = True, scale: Optional[float] = None, window=None, cu_seqlens=None):
    """Fused attention; GQA when ``k.shape[2] < q.shape[2]``.  Returns ``[sq, b, hq, d]``.  ``window = (left, right)``: sliding window; ``cu_seqlens``: the
    token dim holds several packed sequences (THD; ``b == 1``) that must not see each other.  Both run inside the native kernels as a band mask."""
    import math

    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if _use_cuda(q) and _native_attention_ok(q, k, v, causal, window, cu_seqlens):
        band = attention_band(q.shape[0], k.shape[0], window, cu_seqlens, q.device)
        if band is None:
            return _FlashAttnFn.apply(q, k, v, causal, scale)
        return _FlashAttnFn.apply(q, k, v, causal, scale, band[0], band[1])
    if cu_seqlens is not None:
        return ref.attention_fwd(q, k, v, causal, scale, window, cu_seqlens)
    if q.is_cuda:
        # library path (cuDNN/flash SDPA) for shapes the native kernel does not cover
        qb, kb, vb = (t.permute(1, 2, 0, 3) for t in (q, k, v))
        if window is None:
            from torch.nn.attention import SDPBackend, sdpa_kernel

            # cuDNN's Blackwell kernel first (measured 1.1-1.2 PF on B200 vs 0.3 PF for the sm_80 flash
            # kernel); never the O(s^2)-memory math path.
            with sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION], set_priority=True):
                o = torch.nn.functional.scaled_dot_product_attention(
                    qb, kb, vb, is_causal=causal and q.shape[0] == k.shape[0], scale=scale, enable_gqa=kb.shape[1] != qb.shape[1]
                )
            return o.permute(2, 0, 1, 3).contiguous()
    return ref.attention_fwd(q, k, v, causal, scale, window)


# =============================================================================
# Paged-KV decode (inference)
# =============================================================================


def paged_kv_append(k_new, v_new, k_pool, v_pool, block_table, positions) -> None:
    """Write one new K/V entry per request into its page (``k_new/v_new [B, hk, d]``; pools ``[num_blocks, block_size, hk, d]`` of ONE layer;
    ``block_table [B, W]`` / ``positions [B]``).  One CUDA kernel for K and V (``csrc/paged_attention.cu``); index ops on CPU."""
    if _use_cuda(k_new) and hasattr(ext(), "paged_kv_append") and k_new.dtype == torch.bfloat16:
        ext().paged_kv_append(k_new.contiguous(), v_new.contiguous(), k_pool, v_pool, block_table.to(torch.int32).contiguous(), positions.to(torch.int32).contiguous())
        _count()
        return
    bs = k_pool.shape[1]
    pos = positions.long()
    blk = block_table.long().gather(1, (pos // bs).unsqueeze(1)).squeeze(1)
    k_pool[blk, pos % bs] = k_new
    v_pool[blk, pos % bs] = v_new


def paged_attention_decode(q, k_pool, v_pool, block_table, lengths, scale: float, max_len: int):
    """One new query token per request attends to its own paged history: ``q [B, hq, d]`` → ``[B, hq, d]``; ``lengths [B]`` counts the valid tokens INCLUDING
    the one just appended.  CUDA: flash-decoding over the block table (split-KV, GQA heads share every K/V load), no gather; CPU: masked SDPA over a gather."""
    B, hq, d = q.shape
    hk = k_pool.shape[2]
    if _use_cuda(q) and hasattr(ext(), "paged_decode") and q.dtype == torch.bfloat16 and d in (64, 128) and hq % hk == 0 and hq // hk in (1, 2, 4, 8):
        from ..core.transformer.custom_layers.batch_invariant_kernels import is_batch_invariant_mode_enabled

        # batch-invariant mode: split boundaries at fixed positions (512 tokens) so that a request's reduction order does not depend on its batch
        out = ext().paged_decode(q.contiguous(), k_pool, v_pool, block_table.to(torch.int32).contiguous(), lengths.to(torch.int32).contiguous(), float(scale), int(max_len),
                                 512 if is_batch_invariant_mode_enabled() else 0)
        _count(2)
        return out
    bs = k_pool.shape[1]
    nblk = (int(max_len) + bs - 1) // bs
    t = block_table.long()[:, :nblk]
    K = k_pool[t].reshape(B, nblk * bs, hk, d)
    V = v_pool[t].reshape(B, nblk * bs, hk, d)
    rep = hq // hk
    L = nblk * bs
    qf = q.reshape(B * hk, rep, 1, d)
    Kf = K.permute(0, 2, 1, 3).reshape(B * hk, 1, L, d).expand(B * hk, rep, L, d)
    Vf = V.permute(0, 2, 1, 3).reshape(B * hk, 1, L, d).expand(B * hk, rep, L, d)
    mask = (torch.arange(L, device=q.device)[None, :] < lengths.long()[:, None]).repeat_interleave(hk, 0).view(B * hk, 1, 1, L)
    out = torch.nn.functional.scaled_dot_product_attention(qf, Kf, Vf, attn_mask=mask, scale=scale)
    return out.reshape(B, hq, d)


# =============================================================================
# Cross entropy (vocab-parallel, fused)
# =============================================================================


class _VocabParallelCEFn(torch.autograd.Function):
    """Fused vocab-parallel CE: one pass for (max, sum-exp, target logit) per row,
    ONE all-reduce-able stats tensor, and an in-place backward that overwrites the
    logits buffer with ``softmax - onehot`` (reference: 3 ARs, ``cross_entropy.py:130-152``)."""

    @staticmethod
    def forward(ctx, logits, target, group, label_smoothing, vocab_start, reduce_fn):
        shape = logits.shape
        V = shape[-1]
        l2 = logits.reshape(-1, V)
        t1 = target.reshape(-1)
        if _use_cuda(l2):
            l2 = l2.contiguous()
            stats = ext().ce_stats(l2, t1, int(vocab_start))  # [3, rows] fp32: max, sumexp(rel. to max), picked
            _count()
        else:
            lf = l2.float()
            m = lf.max(dim=-1).values
            se = torch.exp(lf - m.unsqueeze(-1)).sum(-1)
            local = t1 - vocab_start
            inr = (local >= 0) & (local < V)
            picked = torch.where(inr, lf.gather(-1, local.clamp(0, V - 1).unsqueeze(-1)).squeeze(-1), torch.zeros_like(m))
            stats = torch.stack([m, se, picked])
        if reduce_fn is not None:
            gmax = reduce_fn(stats[0].clone(), "max")
            se = stats[1] * torch.exp(stats[0] - gmax)
            both = reduce_fn(torch.stack([se, stats[2]]), "sum")
            se, picked = both[0], both[1]
        else:
            gmax, se, picked = stats[0], stats[1], stats[2]
        lse = gmax + torch.log(se)
        loss = lse - picked
        ctx.label_smoothing = label_smoothing
        if label_smoothing > 0:
            # needs mean log-prob over the full vocab
            sum_logits = l2.float().sum(-1)
            nvocab = torch.tensor(float(V), device=l2.device)
            if reduce_fn is not None:
                sum_logits = reduce_fn(sum_logits, "sum")
                nvocab = reduce_fn(nvocab.reshape(1), "sum")[0]
            mean_logprob = sum_logits / nvocab - lse
            smooth = label_smoothing * nvocab / (nvocab - 1)
            loss = (1 - smooth) * loss - smooth * mean_logprob
            ctx.smooth, ctx.nvocab = smooth, nvocab
        ctx.save_for_backward(l2, t1, lse)
        ctx.vocab_start, ctx.shape = vocab_start, shape
        return loss.view(shape[:-1])

    @staticmethod
    def backward(ctx, gloss):
        l2, t1, lse = ctx.saved_tensors
        g = gloss.reshape(-1).float().contiguous()
        V = l2.shape[-1]
        if _use_cuda(l2) and ctx.label_smoothing == 0:
            grad = ext().ce_bwd(l2, t1, lse, g, int(ctx.vocab_start))  # in place on l2's storage
            _count()
        else:
            p = torch.exp(l2.float() - lse.unsqueeze(-1))
            local = t1 - ctx.vocab_start
            inr = (local >= 0) & (local < V)
            onehot = torch.zeros_like(p)
            onehot.scatter_(1, local.clamp(0, V - 1).unsqueeze(-1), inr.float().unsqueeze(-1))
            if ctx.label_smoothing > 0:
                s = ctx.smooth
                grad = p - (1 - s) * onehot - s / ctx.nvocab
            else:
                grad = p - onehot
            grad = (grad * g.unsqueeze(-1)).to(l2.dtype)
        return grad.view(ctx.shape), None, None, None, None, None


def vocab_parallel_cross_entropy(logits, target, group=None, label_smoothing: float = 0.0, vocab_start: int = 0):
    """Per-token CE for logits sharded along the vocab axis over ``group``."""
    import torch.distributed as dist

    reduce_fn = None
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        def reduce_fn(t, op):
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
            return t

    return _VocabParallelCEFn.apply(logits, target, group, label_smoothing, vocab_start, reduce_fn)


# =============================================================================
# multi-tensor optimizer kernels
# =============================================================================


def multi_tensor_l2norm(tensors: List[torch.Tensor]) -> torch.Tensor:
    """sqrt(sum_i ||t_i||²) as a 0-d fp32 tensor on the tensors' device."""
    if tensors and _use_cuda(tensors[0]):
        out = ext().multi_l2norm([t.detach() for t in tensors])
        _count()
        return out
    return ref.l2norm(tensors)


def multi_tensor_scale(tensors: List[torch.Tensor], scale) -> None:
    if not tensors:
        return
    if _use_cuda(tensors[0]) and isinstance(scale, torch.Tensor):
        ext().multi_scale([t.detach() for t in tensors], scale.float().reshape(1))
        _count()
        return
    for t in tensors:
        t.mul_(scale)


def fused_adam(params32, grads, exp_avgs, exp_avg_sqs, lowp_params, *, lr, beta1, beta2, eps, weight_decay, step, adamw=True, grad_scale=None):
    """One launch updates fp32 master, moments and the bf16 model copy for a list of tensors.

    ``grad_scale`` (0-d/1-elem fp32 tensor or None) multiplies the gradient first — it
    carries 1/loss_scale * clip_coef so no separate unscale/clip pass over memory is needed.
    """
    if params32 and _use_cuda(params32[0]):
        gs = grad_scale if grad_scale is not None else torch.ones(1, dtype=torch.float32, device=params32[0].device)
        ext().multi_adam(
            params32, grads, exp_avgs, exp_avg_sqs, [p if p is not None else params32[i] for i, p in enumerate(lowp_params)],
            float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step), bool(adamw), gs.float().reshape(1),
        )
        _count()
        return
    gsv = float(grad_scale) if grad_scale is not None else 1.0
    for p, g, m, v, lp in zip(params32, grads, exp_avgs, exp_avg_sqs, lowp_params):
        ref.adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, adamw, gsv, lp)


# =============================================================================
# MoE: permutation / combination / routing (csrc/moe_kernels.cu)
# =============================================================================


def _moe_native(t: torch.Tensor) -> bool:
    return _use_cuda(t) and hasattr(ext(), "moe_gather_rows") and t.dtype == torch.bfloat16 and t.dim() == 2 and t.shape[1] % 8 == 0


class _GatherRowsFn(torch.autograd.Function):
    """out[i] = scale[i] * x[idx[i]];  backward scatters (index_add) the rows back."""

    @staticmethod
    def forward(ctx, x, idx, scale):
        ctx.save_for_backward(x, idx, scale if scale is not None else x.new_empty(0))
        ctx.has_scale = scale is not None
        _count()
        return ext().moe_gather_rows(x.contiguous(), idx.contiguous(), scale.float().contiguous() if scale is not None else None)

    @staticmethod
    def backward(ctx, g):
        x, idx, scale = ctx.saved_tensors
        g = g.contiguous()
        gs = None
        if ctx.has_scale:
            gs = (g.float() * x.index_select(0, idx).float()).sum(-1).to(scale.dtype)
            g = g * scale.unsqueeze(-1).to(g.dtype)
        gx = torch.zeros_like(x).index_add_(0, idx, g)
        return gx, None, gs


class _CombineRowsFn(torch.autograd.Function):
    """out[t] = Σ_k w[t,k] * x[pos[t,k]]  (pos < 0 skipped); deterministic fp32 accumulation, no atomics."""

    @staticmethod
    def forward(ctx, x, pos, w):
        ctx.save_for_backward(x, pos, w if w is not None else x.new_empty(0))
        ctx.has_w = w is not None
        _count()
        return ext().moe_combine_rows(x.contiguous(), pos.contiguous(), w.float().contiguous() if w is not None else None)

    @staticmethod
    def backward(ctx, g):
        x, pos, w = ctx.saved_tensors
        g = g.contiguous()
        T, K = pos.shape
        valid = pos >= 0
        rows = pos[valid]                                    # permuted rows that received a contribution
        tok = torch.arange(T, device=pos.device).unsqueeze(1).expand(T, K)[valid]
        gx = torch.zeros_like(x)
        if ctx.has_w:
            wv = w[valid].float()
            gx[rows] = ext().moe_gather_rows(g, tok.contiguous(), wv.contiguous())
            gw = torch.zeros_like(w, dtype=torch.float32)
            gw[valid] = (x.index_select(0, rows).float() * g.index_select(0, tok).float()).sum(-1)
            return gx, None, gw.to(w.dtype)
        gx[rows] = ext().moe_gather_rows(g, tok.contiguous(), None)
        return gx, None, None


def moe_gather_rows(x: torch.Tensor, idx: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    if _moe_native(x):
        return _GatherRowsFn.apply(x, idx, scale)
    out = x.index_select(0, idx)
    return out * scale.unsqueeze(-1).to(out.dtype) if scale is not None else out


def moe_combine_rows(x: torch.Tensor, pos: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M, h] permuted rows; pos [T, k] int64 row of x feeding token t (or -1); w [T, k] weights."""
    if _moe_native(x):
        return _CombineRowsFn.apply(x, pos, w)
    g = x.index_select(0, pos.clamp(min=0).reshape(-1)).view(*pos.shape, x.shape[1]).float()
    m = (pos >= 0).unsqueeze(-1).float()
    if w is not None:
        m = m * w.unsqueeze(-1).float()
    return (g * m).sum(1).to(x.dtype)


def moe_topk_router(logits: torch.Tensor, topk: int, score_function: str = "softmax", pre_softmax: bool = False, expert_bias: Optional[torch.Tensor] = None,
                    scaling_factor: Optional[float] = None):
    """Selection part of the router in ONE kernel → (ids [T,k], routing_map [T,E] bool, tokens_per_expert [E] int32, probs [T,k] fp32).
    The returned probs carry no autograd graph; callers that train the router re-derive them from the logits at ``ids``."""
    fn = {"softmax": 0 if pre_softmax else 2, "sigmoid": 1}[score_function]
    if _use_cuda(logits) and hasattr(ext(), "moe_topk_router") and logits.shape[1] <= 256 and topk <= 8:
        _count()
        probs, ids, rmap, tpe = ext().moe_topk_router(logits.detach().float().contiguous(), None if expert_bias is None else expert_bias.float().contiguous(), topk, fn,
                                                      fn == 1, float(scaling_factor or 1.0))
        return ids, rmap, tpe, probs
    lf = logits.detach().float()
    scores = torch.softmax(lf, -1) if fn == 0 else (torch.sigmoid(lf) if fn == 1 else lf)
    key = scores + expert_bias.float() if expert_bias is not None else scores
    ids = torch.topk(key, topk, dim=1).indices
    vals = scores.gather(1, ids)
    probs = torch.softmax(vals, -1) if fn == 2 else (vals / (vals.sum(-1, keepdim=True) + 1e-20) if (fn == 1 and topk > 1) else vals)
    rmap = torch.zeros_like(lf, dtype=torch.bool).scatter(1, ids, True)
    return ids, rmap, rmap.sum(0).int(), probs * float(scaling_factor or 1.0)


# second kernel batch: residual-fused RMSNorm, RoPE thd / fused-QKV, causal conv1d, SSD state passing / step, MXFP8 (see ops/extra.py)
from .extra import (  # noqa: E402,F401
    add_rms_norm,
    apply_rope_qkv,
    apply_rope_thd,
    causal_conv1d,
    gemm_mxfp8_nt,
    gemm_nvfp4_nt,
    mxfp8_dequantize,
    mxfp8_quantize,
    mxfp8_quantize_reference,
    mxfp8_swizzle_scales,
    nvfp4_pack,
    nvfp4_quantize,
    nvfp4_unpack,
    positions_from_cu_seqlens,
    quick_geglu,
    scaled_masked_softmax,
    squared_relu,
    ssd_state_passing,
    ssd_step,
)
