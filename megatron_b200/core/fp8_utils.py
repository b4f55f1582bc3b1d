"""FP8 training support (reference ``core/fp8_utils.py`` 1,056 LoC + TE quantisers).

Recipes implemented natively:
* ``tensorwise`` (current scaling): scale = fp8_max / amax(tensor), computed on the GPU right before the cast — no host sync;
  the dequantisation scale stays a device scalar that the tcgen05 FP8 GEMM epilogue multiplies in.
* ``delayed``: scale from the running max of an amax history (``Fp8Meta``), updated after each use.
Forward operands are E4M3; gradients E5M2 (``hybrid`` format) or E4M3.

``fp8_linear(x, w)`` is the autograd op used by ``ColumnParallelLinear``/``RowParallelLinear`` when ``config.fp8`` is set: all three
GEMMs (fwd, dgrad, wgrad) run on ``ops.ext().gemm_fp8_nt`` (``csrc/gemm_fp8_sm100.cu``, ``tcgen05.mma kind::f8f6f4``) through K-major
("NT") operand copies, the same cast+transpose data flow TransformerEngine uses.  Without the CUDA extension the math falls back to
emulated FP8 (cast → fp32 matmul) so numerics tests run on CPU.
"""
from __future__ import annotations

from contextlib import contextmanager, nullcontext
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

E4M3, E5M2 = torch.float8_e4m3fn, torch.float8_e5m2
FP8_MAX = {E4M3: 448.0, E5M2: 57344.0}


def is_float8tensor(t) -> bool:
    return isinstance(t, torch.Tensor) and t.dtype in (E4M3, E5M2)


@dataclass
class Fp8Meta:
    """Delayed-scaling state of one tensor role (input / weight / grad)."""

    history_len: int = 16
    margin: int = 0
    amax_history: Optional[torch.Tensor] = None
    scale: Optional[torch.Tensor] = None

    def scale_for(self, t: torch.Tensor, dtype) -> torch.Tensor:
        amax_now = t.detach().abs().amax().float()
        if self.amax_history is None:
            self.amax_history = amax_now.repeat(self.history_len)
            amax = amax_now
        else:
            amax = self.amax_history.max()
            self.amax_history = torch.roll(self.amax_history, 1)
            self.amax_history[0] = amax_now
        self.scale = (FP8_MAX[dtype] / (2.0**self.margin)) / amax.clamp(min=1e-12)
        return self.scale


def quantize(t: torch.Tensor, dtype=E4M3, meta: Optional[Fp8Meta] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """→ (fp8 tensor, dequantisation scale as a 1-element fp32 tensor on the same device)."""
    if meta is not None:
        scale = meta.scale_for(t, dtype)
    else:
        scale = FP8_MAX[dtype] / t.detach().abs().amax().float().clamp(min=1e-12)
    q = (t.float() * scale).clamp(-FP8_MAX[dtype], FP8_MAX[dtype]).to(dtype)
    return q, (1.0 / scale).reshape(1)


def _gemm_nt(aq, a_inv, bq, b_inv) -> torch.Tensor:
    """bf16 [M, N] = (aq · bqᵀ) * a_inv * b_inv."""
    from .. import ops

    if aq.is_cuda and ops.has_ext() and hasattr(ops.ext(), "gemm_fp8_nt") and aq.shape[1] % 16 == 0 and bq.shape[0] % 8 == 0:
        ops._count()
        return ops.ext().gemm_fp8_nt(aq.contiguous(), bq.contiguous(), 1.0, (a_inv * b_inv).float())
    return ((aq.float() @ bq.float().t()) * (a_inv * b_inv)).to(torch.bfloat16)


class _Fp8LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, grad_dtype, metas):
        x2 = x.reshape(-1, x.shape[-1])
        xq, xi = quantize(x2, E4M3, metas[0] if metas else None)
        wq, wi = quantize(w, E4M3, metas[1] if metas else None)
        ctx.save_for_backward(xq, xi, wq, wi)
        ctx.grad_dtype, ctx.metas, ctx.x_shape, ctx.out_dtype = grad_dtype, metas, x.shape, x.dtype
        return _gemm_nt(xq, xi, wq, wi).to(x.dtype).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        xq, xi, wq, wi = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        gq, gi = quantize(g2, ctx.grad_dtype, ctx.metas[2] if ctx.metas else None)
        # dgrad [M, K] = gy [M, N] · W [N, K]      → NT with Wᵀ [K, N]
        gx = _gemm_nt(gq, gi, wq.t().contiguous(), wi).to(ctx.out_dtype).view(ctx.x_shape)
        # wgrad [N, K] = gyᵀ [N, M] · x [M, K]     → NT with gyᵀ [N, M] and xᵀ [K, M]
        gw = _gemm_nt(gq.t().contiguous(), gi, xq.t().contiguous(), xi).to(ctx.out_dtype)
        return gx, gw, None, None


def _pad_rows(t: torch.Tensor, mult: int) -> torch.Tensor:
    r = (-t.shape[0]) % mult
    return t if r == 0 else torch.cat([t, t.new_zeros(r, t.shape[1])], dim=0)


def _mx_gemm_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 [M, N] = a [M, K] · b [N, K]ᵀ with both operands quantised to MXFP8 along K (1x32 blocks, E8M0 scales applied inside the tensor core)."""
    from .. import ops

    K = a.shape[1]
    if K % 128:                      # zero padding along K leaves the product unchanged
        pad = (-K) % 128
        a, b = torch.nn.functional.pad(a, (0, pad)), torch.nn.functional.pad(b, (0, pad))
    aq, asf = ops.mxfp8_quantize(a.to(torch.bfloat16))
    bq, bsf = ops.mxfp8_quantize(b.to(torch.bfloat16))
    return ops.gemm_mxfp8_nt(aq, asf, bq, bsf)


class _Mxfp8SPColumnFn(torch.autograd.Function):
    """Sequence-parallel column linear with the MXFP8 recipe and a QUANTISED all-gather: the activation shard is quantised row-wise before it goes on the
    wire (1.03 bytes/element instead of 2) and the gathered payload feeds the block-scaled GEMM as is.  Backward re-gathers in bf16 (the wgrad GEMM
    reduces over tokens and needs column-wise blocks) and reduce-scatters dgrad, like the bf16 path."""

    @staticmethod
    def forward(ctx, x, w, group):
        from .. import ops
        from ..parallel.quantized_collectives import all_gather_mxfp8

        s_local = x.shape[0]
        x2 = x.reshape(-1, x.shape[-1])                               # [s/tp · b, K] — rows of one rank are contiguous in the gathered [s · b, K]
        K = x2.shape[1]
        assert K % 128 == 0, "quantised all-gather needs hidden % 128 == 0"
        xq, xsf = all_gather_mxfp8(x2, group)
        wq, wsf = ops.mxfp8_quantize(w.to(torch.bfloat16))
        y = ops.gemm_mxfp8_nt(xq, xsf, wq, wsf)
        ctx.save_for_backward(x, w)
        ctx.group = group
        ws = xq.shape[0] // x2.shape[0]
        return y.to(x.dtype).view(s_local * ws, *x.shape[1:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        import torch.distributed as dist

        x, w = ctx.saved_tensors
        ws = dist.get_world_size(ctx.group)
        full = x.new_empty((x.shape[0] * ws,) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(full, x.contiguous(), group=ctx.group)
        g2 = gy.reshape(-1, gy.shape[-1])
        gx_full = _mx_gemm_nt(g2, w.t().contiguous()).to(x.dtype).view(full.shape)
        gx = x.new_empty(x.shape)
        dist.reduce_scatter_tensor(gx, gx_full.contiguous(), group=ctx.group)
        gw = _mx_gemm_nt(g2.t().contiguous(), full.reshape(-1, full.shape[-1]).t().contiguous()).to(w.dtype)
        return gx, gw, None


def mxfp8_sp_column_linear(x: torch.Tensor, w: torch.Tensor, group) -> torch.Tensor:
    return _Mxfp8SPColumnFn.apply(x, w, group)


class _Mxfp8LinearFn(torch.autograd.Function):
    """MXFP8 recipe (reference ``fp8_recipe="mxfp8"``): every GEMM operand is quantised along ITS reduction dimension, so the backward GEMMs
    re-quantise transposed copies (row-wise for fprop, column-wise for dgrad / wgrad) — the data flow of TE's MXFP8 tensors with both usages."""

    @staticmethod
    def forward(ctx, x, w):
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, w)
        ctx.x_shape = x.shape
        return _mx_gemm_nt(x2, w).to(x.dtype).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        gx = _mx_gemm_nt(g2, w.t().contiguous()).to(x2.dtype).view(ctx.x_shape)          # [M, N] · [K, N]ᵀ, reduction over N
        gw = _mx_gemm_nt(g2.t().contiguous(), x2.t().contiguous()).to(w.dtype)           # [N, M] · [K, M]ᵀ, reduction over M
        return gx, gw


def _nvf4_gemm_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 [M, N] = a [M, K] · b [N, K]ᵀ with both operands quantised to NVFP4 along K (16-element blocks, UE4M3 block scales, fp32 tensor scale)."""
    from .. import ops

    K = a.shape[1]
    if K % 256:
        pad = (-K) % 256
        a, b = torch.nn.functional.pad(a, (0, pad)), torch.nn.functional.pad(b, (0, pad))
    return ops.gemm_nvfp4_nt(*ops.nvfp4_quantize(a), *ops.nvfp4_quantize(b))


def _nvf4_gemm_nt_sr(a: torch.Tensor, b: torch.Tensor, a_stochastic: bool, rht: bool) -> torch.Tensor:
    """NVFP4 GEMM for the backward pass: optional random-Hadamard-free deterministic 16-point Hadamard on the reduction dim of both operands (``rht``) and
    stochastic rounding of the gradient operand ``a``."""
    from .. import ops
    from .fp4_utils import hadamard16, quantize_nvfp4

    K = a.shape[1]
    pad = (-K) % 256
    if pad:
        a, b = torch.nn.functional.pad(a, (0, pad)), torch.nn.functional.pad(b, (0, pad))
    if rht:
        a, b = hadamard16(a), hadamard16(b)
    qa = quantize_nvfp4(a, stochastic=True) if a_stochastic else ops.nvfp4_quantize(a)
    if a_stochastic:
        qa = (ops.nvfp4_pack(qa[0]), qa[1], qa[2].reshape(1))
    return ops.gemm_nvfp4_nt(*qa, *ops.nvfp4_quantize(b))


class _Nvfp4LinearFn(torch.autograd.Function):
    """NVFP4 recipe (reference ``fp4_recipe="nvfp4"``).  Forward GEMM in 4 bits (nearest rounding).  ``full_fp4`` runs dgrad and wgrad in 4 bits as well, with
    the two ingredients that make 4-bit gradients trainable: stochastic rounding of the gradient operand (unbiased) and a 16-point Hadamard rotation along
    the reduction dimension of the wgrad GEMM (outlier spreading; cancels in the product).  Without ``full_fp4`` the backward GEMMs use MXFP8."""

    @staticmethod
    def forward(ctx, x, w, full_fp4):
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, w)
        ctx.x_shape, ctx.full_fp4 = x.shape, full_fp4
        return _nvf4_gemm_nt(x2, w).to(x.dtype).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.full_fp4:
            gx = _nvf4_gemm_nt_sr(g2, w.t().contiguous(), a_stochastic=True, rht=False).to(x2.dtype).view(ctx.x_shape)
            gw = _nvf4_gemm_nt_sr(g2.t().contiguous(), x2.t().contiguous(), a_stochastic=True, rht=True).to(w.dtype)
        else:
            gx = _mx_gemm_nt(g2, w.t().contiguous()).to(x2.dtype).view(ctx.x_shape)
            gw = _mx_gemm_nt(g2.t().contiguous(), x2.t().contiguous()).to(w.dtype)
        return gx, gw, None


def fp8_linear(x: torch.Tensor, w: torch.Tensor, recipe: str = "tensorwise", fp8_format: str = "hybrid", metas=None) -> torch.Tensor:
    """``x [..., K] @ w [N, K]ᵀ`` with FP8 operands for all three GEMMs of the layer.  ``recipe``: ``tensorwise`` | ``delayed`` | ``mxfp8`` | ``nvfp4`` (4-bit forward, MXFP8 backward)."""
    if recipe == "mxfp8":
        return _Mxfp8LinearFn.apply(x, w)
    if recipe in ("nvfp4", "nvfp4_full"):
        return _Nvfp4LinearFn.apply(x, w, recipe == "nvfp4_full")
    grad_dtype = E5M2 if fp8_format == "hybrid" else E4M3
    if recipe == "delayed" and metas is None:
        raise ValueError("delayed scaling needs (input, weight, grad) Fp8Meta objects")
    return _Fp8LinearFn.apply(x, w, grad_dtype, metas if recipe == "delayed" else None)


class Fp8LinearState(torch.nn.Module):
    """Holder of the three delayed-scaling metas of one linear layer (checkpointed as extra state)."""

    def __init__(self, history_len: int = 16, margin: int = 0):
        super().__init__()
        self.metas = (Fp8Meta(history_len, margin), Fp8Meta(history_len, margin), Fp8Meta(history_len, margin))

    def get_extra_state(self):
        return [{"amax_history": m.amax_history, "scale": m.scale} for m in self.metas]

    def set_extra_state(self, state):
        for m, s in zip(self.metas, state or []):
            m.amax_history, m.scale = s.get("amax_history"), s.get("scale")


_FP8_ENABLED = [False]


def fp8_enabled() -> bool:
    return _FP8_ENABLED[-1]


@contextmanager
def fp8_autocast(enabled: bool = True):
    _FP8_ENABLED.append(bool(enabled))
    try:
        yield
    finally:
        _FP8_ENABLED.pop()


def get_fp8_context(config, layer_no: int = -1, is_init: bool = False):
    """Per-layer FP8 context (reference ``fp8_utils.py:832``): the first/last ``num_layers_at_{start,end}_in_bf16`` layers stay bf16."""
    if not getattr(config, "fp8", None):
        return nullcontext()
    n = config.num_layers
    lo = getattr(config, "num_layers_at_start_in_bf16", 0) if getattr(config, "first_last_layers_bf16", False) else 0
    hi = getattr(config, "num_layers_at_end_in_bf16", 0) if getattr(config, "first_last_layers_bf16", False) else 0
    if 0 <= layer_no < lo or (layer_no >= 0 and layer_no >= n - hi):
        return fp8_autocast(False)
    return fp8_autocast(True)
