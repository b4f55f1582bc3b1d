"""top-k indices ↔ multi-hot routing map (reference ``fusions/fused_indices_converter.py``, 2 Triton kernels; used by DeepEP-style dispatch)."""
import torch


def fused_indices_to_multihot(indices: torch.Tensor, probs_indices: torch.Tensor, num_of_local_experts: int):
    """indices [T, k] (−1 = dropped), probs [T, k] → (multihot [T, E] bool, probs [T, E])."""
    T, k = indices.shape
    valid = indices >= 0
    idx = indices.clamp(min=0)
    multihot = torch.zeros(T, num_of_local_experts, dtype=torch.bool, device=indices.device)
    multihot.scatter_(1, idx, valid)
    probs = torch.zeros(T, num_of_local_experts, dtype=probs_indices.dtype, device=indices.device)
    probs.scatter_add_(1, idx, probs_indices * valid.to(probs_indices.dtype))
    return multihot, probs


def fused_multihot_to_indices(multihot: torch.Tensor, probs: torch.Tensor, topk: int):
    """Inverse: routing map [T, E] → (indices [T, k] padded with −1, probs [T, k])."""
    vals, idx = torch.topk(multihot.to(torch.int8), topk, dim=1)
    idx = torch.where(vals > 0, idx, torch.full_like(idx, -1))
    return idx, torch.where(vals > 0, probs.gather(1, idx.clamp(min=0)), torch.zeros_like(probs[:, :topk]))
