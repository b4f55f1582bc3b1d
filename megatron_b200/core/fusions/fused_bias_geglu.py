"""bias + GeGLU and (token-weighted, clamped, offset) quick-GeGLU (reference ``fusions/fused_bias_geglu.py:16-442``).

Hot paths: ``ops.geglu`` / ``ops.extra.quick_geglu`` (CUDA).  Weighted / clamped / offset variants (GPT-OSS MoE:
``alpha = 1.702``, ``clamp = 7``, ``offset = 1``) and the explicit backward functions: ``_gated.py``."""
import torch

from ... import ops
from ._gated import gated_activation, gated_backward, gated_forward


def quick_gelu(y: torch.Tensor, alpha: float = 1.702) -> torch.Tensor:
    return y * torch.sigmoid(alpha * y)


def geglu(y):
    return ops.geglu(y)


def bias_geglu(bias, y):
    return ops.geglu(y, bias)


def geglu_back(g, y):
    return gated_backward(g, y, None, None, "gelu")[0]


def bias_geglu_back(g, y, bias):
    return gated_backward(g, y, bias, None, "gelu")[0]


def quick_geglu(y, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    return gated_forward(y, None, None, "quick_gelu", clamp_value, linear_offset, alpha)


def weighted_quick_geglu(y, weights, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    return gated_forward(y, None, weights, "quick_gelu", clamp_value, linear_offset, alpha)


def weighted_bias_quick_geglu(y, bias, weights, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    return gated_forward(y, bias, weights, "quick_gelu", clamp_value, linear_offset, alpha)


def quick_geglu_back(g, y, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    return gated_backward(g, y, None, None, "quick_gelu", clamp_value, linear_offset, alpha)[0]


def weighted_quick_geglu_back(g, y, weights, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    gy, _, gw = gated_backward(g, y, None, weights, "quick_gelu", clamp_value, linear_offset, alpha)
    return gy, gw


def weighted_bias_quick_geglu_back(g, y, bias, weights, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    return gated_backward(g, y, bias, weights, "quick_gelu", clamp_value, linear_offset, alpha)


class GeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return gated_forward(input, None, None, "gelu")

    @staticmethod
    def backward(ctx, g):
        return gated_backward(g, ctx.saved_tensors[0], None, None, "gelu")[0]


class BiasGeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias):
        ctx.save_for_backward(input, bias)
        return gated_forward(input, bias, None, "gelu")

    @staticmethod
    def backward(ctx, g):
        y, b = ctx.saved_tensors
        gy, gb, _ = gated_backward(g, y, b, None, "gelu")
        return gy, gb


class WeightedQuickGeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weights, fp8_input_store=False, linear_offset=0.0, clamp_value=None, alpha=1.702):
        ctx.cfg = (linear_offset, clamp_value, alpha, fp8_input_store, input.dtype)
        ctx.save_for_backward(input.to(torch.float8_e4m3fn) if fp8_input_store else input, weights)
        return gated_forward(input, None, weights, "quick_gelu", clamp_value, linear_offset, alpha)

    @staticmethod
    def backward(ctx, g):
        y, w = ctx.saved_tensors
        off, clamp, alpha, fp8, dtype = ctx.cfg
        gy, _, gw = gated_backward(g, y.to(dtype) if fp8 else y, None, w, "quick_gelu", clamp, off, alpha)
        return gy, gw, None, None, None, None


class WeightedBiasQuickGeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias, weights, fp8_input_store=False, linear_offset=0.0, clamp_value=None, alpha=1.702):
        ctx.cfg = (linear_offset, clamp_value, alpha, fp8_input_store, input.dtype)
        ctx.save_for_backward(input.to(torch.float8_e4m3fn) if fp8_input_store else input, bias, weights)
        return gated_forward(input, bias, weights, "quick_gelu", clamp_value, linear_offset, alpha)

    @staticmethod
    def backward(ctx, g):
        y, b, w = ctx.saved_tensors
        off, clamp, alpha, fp8, dtype = ctx.cfg
        gy, gb, gw = gated_backward(g, y.to(dtype) if fp8 else y, b, w, "quick_gelu", clamp, off, alpha)
        return gy, gb, gw, None, None, None, None


def bias_geglu_impl(input, bias):
    shape = input.shape
    y = ops.geglu(input.reshape(-1, shape[-1]), bias)
    return y.view(*shape[:-1], shape[-1] // 2)


def weighted_bias_quick_geglu_impl(input, bias, weights, fp8_input_store=False, linear_offset: float = 0.0, clamp_value=None, alpha: float = 1.702):
    """Token-weighted quick-GeGLU of the MoE experts: one autograd node with a recomputing fp32 backward."""
    shape = input.shape
    x = input.reshape(-1, shape[-1])
    w = weights.reshape(-1, 1)
    if bias is not None:
        y = WeightedBiasQuickGeGLUFunction.apply(x, bias, w, fp8_input_store, linear_offset, clamp_value, alpha)
    else:
        y = WeightedQuickGeGLUFunction.apply(x, w, fp8_input_store, linear_offset, clamp_value, alpha)
    return y.view(*shape[:-1], shape[-1] // 2)
