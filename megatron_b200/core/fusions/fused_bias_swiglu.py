"""bias + SwiGLU, token-weighted and clamped variants (reference ``fusions/fused_bias_swiglu.py:15-341``).

Hot path (no clamp): ``ops.swiglu`` — one CUDA kernel forward, one backward (``csrc/elementwise.cu``, 4.3–5.7 TB/s).
Clamped variants (GPT-OSS) and the explicit ``*_back`` functions: ``_gated.py``."""
import torch

from ... import ops
from ._gated import GatedActivationFunction, gated_activation, gated_backward, gated_forward


# ---- forward / backward as plain functions (the reference's jit-fused bodies) -------------------------------------------
def swiglu(y):
    return ops.swiglu(y)


def bias_swiglu(y, bias):
    return ops.swiglu(y, bias)


def weighted_swiglu(y, weights):
    return gated_forward(y, None, weights, "silu")


def clamped_swiglu(y, clamp_value):
    return gated_forward(y, None, None, "silu", clamp_value)


def bias_clamped_swiglu(y, bias, clamp_value):
    return gated_forward(y, bias, None, "silu", clamp_value)


def clamped_weighted_swiglu(y, weights, clamp_value):
    return gated_forward(y, None, weights, "silu", clamp_value)


def swiglu_back(g, y):
    return gated_backward(g, y, None, None, "silu")[0]


def bias_swiglu_back(g, y, bias):
    return gated_backward(g, y, bias, None, "silu")[0]


def weighted_swiglu_back(g, y, weights):
    gy, _, gw = gated_backward(g, y, None, weights, "silu")
    return gy, gw


def clamped_swiglu_back(g, y, clamp_value):
    return gated_backward(g, y, None, None, "silu", clamp_value)[0]


def bias_clamped_swiglu_back(g, y, bias, clamp_value):
    return gated_backward(g, y, bias, None, "silu", clamp_value)[0]


def clamped_weighted_swiglu_back(g, y, weights, clamp_value):
    gy, _, gw = gated_backward(g, y, None, weights, "silu", clamp_value)
    return gy, gw


# ---- autograd functions ------------------------------------------------------------------------------------------------
class BiasSwiGLUFunction(torch.autograd.Function):
    """``apply(input, bias, fp8_input_store, cpu_offload_input, clamp_value=None)``"""

    @staticmethod
    def forward(ctx, input, bias, fp8_input_store=False, cpu_offload_input=False, clamp_value=None):
        ctx.cfg = (clamp_value, fp8_input_store, input.dtype)
        saved = input.to(torch.float8_e4m3fn) if fp8_input_store else input
        if cpu_offload_input:
            saved.activation_offloading = True
            bias.activation_offloading = True
        ctx.save_for_backward(saved, bias)
        return gated_forward(input, bias, None, "silu", clamp_value)

    @staticmethod
    def backward(ctx, g):
        y, bias = ctx.saved_tensors
        clamp, fp8, dtype = ctx.cfg
        gy, gb, _ = gated_backward(g, y.to(dtype) if fp8 else y, bias, None, "silu", clamp)
        return gy, gb, None, None, None


class SwiGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, fp8_input_store=False, cpu_offload_input=False, clamp_value=None):
        ctx.cfg = (clamp_value, fp8_input_store, input.dtype)
        saved = input.to(torch.float8_e4m3fn) if fp8_input_store else input
        if cpu_offload_input:
            saved.activation_offloading = True
        ctx.save_for_backward(saved)
        return gated_forward(input, None, None, "silu", clamp_value)

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        clamp, fp8, dtype = ctx.cfg
        return gated_backward(g, y.to(dtype) if fp8 else y, None, None, "silu", clamp)[0], None, None, None


class WeightedSwiGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weights, fp8_input_store=False, clamp_value=None):
        ctx.cfg = (clamp_value, fp8_input_store, input.dtype)
        ctx.save_for_backward(input.to(torch.float8_e4m3fn) if fp8_input_store else input, weights)
        return gated_forward(input, None, weights, "silu", clamp_value)

    @staticmethod
    def backward(ctx, g):
        y, w = ctx.saved_tensors
        clamp, fp8, dtype = ctx.cfg
        gy, _, gw = gated_backward(g, y.to(dtype) if fp8 else y, None, w, "silu", clamp)
        return gy, gw, None, None


# ---- module-facing entry points ----------------------------------------------------------------------------------------------
def bias_swiglu_impl(input, bias, fp8_input_store: bool = False, cpu_offload_input: bool = False, clamp_value=None):
    shape = input.shape
    x = input.reshape(-1, shape[-1])
    if clamp_value is None and not fp8_input_store and not cpu_offload_input:
        y = ops.swiglu(x, bias)                              # CUDA kernel
    elif bias is not None:
        y = BiasSwiGLUFunction.apply(x, bias, fp8_input_store, cpu_offload_input, clamp_value)
    else:
        y = SwiGLUFunction.apply(x, fp8_input_store, cpu_offload_input, clamp_value)
    return y.view(*shape[:-1], shape[-1] // 2)


def weighted_bias_swiglu_impl(input, bias, weights, fp8_input_store: bool = False, clamp_value=None):
    """MoE: ``silu(y1) * y2 * w`` with per-token routing weights ``w [tokens, 1]``."""
    shape = input.shape
    x = input.reshape(-1, shape[-1])
    if clamp_value is None and not fp8_input_store:
        y = ops.swiglu(x, bias, probs=weights.reshape(-1))
    else:
        y = gated_activation(x, bias, weights.reshape(-1, 1), "silu", clamp_value, fp8_input_store=fp8_input_store)
    return y.view(*shape[:-1], shape[-1] // 2)
