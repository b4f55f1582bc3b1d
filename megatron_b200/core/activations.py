"""Activation functions referenced by configs (reference ``core/activations.py``)."""
import torch
import torch.nn.functional as F


def squared_relu(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        from .. import ops

        return ops.squared_relu(x)      # one vectorised kernel fwd, one bwd (csrc/extra_kernels.cu)
    return torch.pow(F.relu(x), 2)


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def fast_gelu(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))
