"""One implementation of every gated activation the MLPs use: ``act(a) * (b + offset) [* w]`` with optional clamping.

``a, b = chunk(y + bias)``; ``act`` is SiLU (SwiGLU), tanh-GELU (GeGLU) or ``a * sigmoid(alpha a)`` (quick-GeGLU, alpha = 1.702).
Clamping follows GPT-OSS: ``a <- min(a, c)``, ``b <- clip(b, -c, c)`` — the gradient is zero where a value was clamped.
The un-clamped, un-offset SiLU / GELU cases go to the CUDA kernels in ``ops`` (``csrc/elementwise.cu``, ``extra_kernels.cu``);
everything else is computed here in fp32 with a hand-derived backward that recomputes from the saved PRE-activation input
(so the saved tensor is the GEMM output that autograd keeps anyway, optionally stored as fp8)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

_K = math.sqrt(2.0 / math.pi)


def _act_and_grad(a: torch.Tensor, kind: str, alpha: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(act(a), d act / d a) in fp32."""
    if kind == "silu":
        s = torch.sigmoid(a)
        return a * s, s * (1 + a * (1 - s))
    if kind == "quick_gelu":
        s = torch.sigmoid(alpha * a)
        return a * s, s * (1 + alpha * a * (1 - s))
    if kind == "gelu":                                       # tanh approximation
        u = _K * (a + 0.044715 * a ** 3)
        t = torch.tanh(u)
        return 0.5 * a * (1 + t), 0.5 * (1 + t) + 0.5 * a * (1 - t * t) * _K * (1 + 3 * 0.044715 * a * a)
    raise ValueError(kind)


def gated_forward(y: torch.Tensor, bias: Optional[torch.Tensor], weights: Optional[torch.Tensor], kind: str, clamp: Optional[float] = None,
                  offset: float = 0.0, alpha: float = 1.702) -> torch.Tensor:
    x = y.float() if bias is None else y.float() + bias.float()
    a, b = torch.chunk(x, 2, dim=-1)
    if clamp is not None:
        a, b = a.clamp(max=clamp), b.clamp(min=-clamp, max=clamp)
    out = _act_and_grad(a, kind, alpha)[0] * (b + offset)
    if weights is not None:
        out = out * weights.float()
    return out.to(y.dtype)


def gated_backward(g: torch.Tensor, y: torch.Tensor, bias: Optional[torch.Tensor], weights: Optional[torch.Tensor], kind: str, clamp: Optional[float] = None,
                   offset: float = 0.0, alpha: float = 1.702):
    """Returns (grad_y, grad_bias or None, grad_weights or None)."""
    x = y.float() if bias is None else y.float() + bias.float()
    a, b = torch.chunk(x, 2, dim=-1)
    gf = g.float()
    if clamp is not None:
        ma, mb = (a <= clamp), (b >= -clamp) & (b <= clamp)
        a, b = a.clamp(max=clamp), b.clamp(min=-clamp, max=clamp)
    act, dact = _act_and_grad(a, kind, alpha)
    gw = None
    if weights is not None:
        gw = (act * (b + offset) * gf).sum(dim=-1, keepdim=True).to(weights.dtype)
        gf = gf * weights.float()
    ga, gb = gf * dact * (b + offset), gf * act
    if clamp is not None:
        ga, gb = ga * ma, gb * mb
    gy = torch.cat([ga, gb], dim=-1)
    gbias = gy.reshape(-1, gy.shape[-1]).sum(0).to(bias.dtype) if bias is not None else None
    return gy.to(y.dtype), gbias, gw


class GatedActivationFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, bias, weights, kind, clamp, offset, alpha, fp8_input_store):
        ctx.cfg = (kind, clamp, offset, alpha, y.dtype, fp8_input_store)
        saved = y.to(torch.float8_e4m3fn) if fp8_input_store else y
        ctx.save_for_backward(saved, bias, weights)
        return gated_forward(y, bias, weights, kind, clamp, offset, alpha)

    @staticmethod
    def backward(ctx, g):
        y, bias, weights = ctx.saved_tensors
        kind, clamp, offset, alpha, dtype, fp8 = ctx.cfg
        gy, gb, gw = gated_backward(g, y.to(dtype) if fp8 else y, bias, weights, kind, clamp, offset, alpha)
        return gy, gb, gw, None, None, None, None, None


def gated_activation(y, bias=None, weights=None, kind="silu", clamp=None, offset=0.0, alpha=1.702, fp8_input_store=False):
    return GatedActivationFunction.apply(y, bias, weights, kind, clamp, offset, alpha, fp8_input_store)
