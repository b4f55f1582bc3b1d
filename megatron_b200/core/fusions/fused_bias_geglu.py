"""bias+GeGLU / quick-GeGLU (reference ``fusions/fused_bias_geglu.py:16-442``)."""
import torch

from ... import ops


def bias_geglu_impl(input, bias):
    shape = input.shape
    y = ops.geglu(input.reshape(-1, shape[-1]), bias)
    return y.view(*shape[:-1], shape[-1] // 2)


def quick_gelu(y: torch.Tensor) -> torch.Tensor:
    return y * torch.sigmoid(1.702 * y)


def weighted_bias_quick_geglu_impl(input, bias, weights, fp8_input_store=False, linear_offset: float = 0.0, clamp_value=None):
    x = input if bias is None else input + bias
    a, b = torch.chunk(x, 2, dim=-1)
    if clamp_value is not None:
        a, b = a.clamp(max=clamp_value), b.clamp(min=-clamp_value, max=clamp_value)
    return (quick_gelu(a) * (b + linear_offset) * weights).to(input.dtype)
