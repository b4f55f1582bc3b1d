"""bias + dropout + residual add (reference ``fusions/fused_bias_dropout.py:11-94``)."""
from ... import ops


def get_bias_dropout_add(training: bool, fused: bool):
    def f(x_with_bias, residual, prob):
        x, bias = x_with_bias
        return ops.bias_dropout_add(x, bias, residual, prob, training)

    return f
