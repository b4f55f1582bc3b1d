"""bias+SwiGLU (reference ``fusions/fused_bias_swiglu.py:15-341``) → ``ops.swiglu`` (``csrc/elementwise.cu``)."""
from ... import ops


def bias_swiglu_impl(input, bias, fp8_input_store: bool = False, cpu_offload_input: bool = False):
    shape = input.shape
    y = ops.swiglu(input.reshape(-1, shape[-1]), bias)
    return y.view(*shape[:-1], shape[-1] // 2)


def weighted_bias_swiglu_impl(input, bias, weights, fp8_input_store: bool = False):
    """MoE: ``silu(y1) * y2 * w`` with per-token routing weights ``w [tokens, 1]``."""
    shape = input.shape
    y = ops.swiglu(input.reshape(-1, shape[-1]), bias, probs=weights.reshape(-1))
    return y.view(*shape[:-1], shape[-1] // 2)


swiglu = lambda y: ops.swiglu(y)  # noqa: E731
bias_swiglu = lambda y, b: ops.swiglu(y, b)  # noqa: E731
