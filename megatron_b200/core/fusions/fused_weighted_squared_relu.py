"""``relu(x)^2 * w`` for MoE experts with squared-ReLU activation (reference ``fusions/fused_weighted_squared_relu.py:13-110``)."""
import torch


def weighted_squared_relu_impl(input: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    r = torch.relu(input.float())
    return (r * r * weights.float()).to(input.dtype)
