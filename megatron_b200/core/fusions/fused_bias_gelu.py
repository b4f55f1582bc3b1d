"""bias+GeLU (tanh approximation; reference ``fusions/fused_bias_gelu.py:16-55``)."""
from ... import ops


def bias_gelu_impl(input, bias):
    return ops.bias_gelu(input, bias)
