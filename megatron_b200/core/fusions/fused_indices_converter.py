"""top-k indices ↔ multi-hot routing map (reference ``fusions/fused_indices_converter.py``: two Triton kernels; used by DeepEP-style dispatch).

CUDA: one warp per token, ``ops/csrc/routing_kernels.cu`` (scatter without pre-zeroed outputs; ordered ballot compaction for the inverse); both directions are
differentiable in the probabilities.  CPU: index ops."""
import torch

from ... import ops


def _use_kernels(*ts) -> bool:
    return all(t.is_cuda for t in ts) and ops.has_ext() and hasattr(ops.ext(), "indices_to_multihot")


class _IndicesToMultihot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indices, probs_indices, num_experts):
        idx = indices.long().contiguous()
        multihot, probs = ops.ext().indices_to_multihot(idx, probs_indices.float().contiguous(), num_experts)
        ops._count()
        ctx.save_for_backward(idx)
        ctx.num_experts, ctx.dtype = num_experts, probs_indices.dtype
        ctx.mark_non_differentiable(multihot)
        return multihot, probs.to(probs_indices.dtype)

    @staticmethod
    def backward(ctx, _g_map, g_probs):
        (idx,) = ctx.saved_tensors
        g = ops.ext().multihot_probs_grad(idx, g_probs.float().contiguous(), ctx.num_experts, False)
        ops._count()
        return None, g.to(ctx.dtype), None


class _MultihotToIndices(torch.autograd.Function):
    @staticmethod
    def forward(ctx, multihot, probs, topk):
        idx, p = ops.ext().multihot_to_indices(multihot.bool().contiguous(), probs.float().contiguous(), topk)
        ops._count()
        ctx.save_for_backward(idx)
        ctx.num_experts, ctx.dtype = multihot.shape[1], probs.dtype
        ctx.mark_non_differentiable(idx)
        return idx, p.to(probs.dtype)

    @staticmethod
    def backward(ctx, _g_idx, g_probs):
        (idx,) = ctx.saved_tensors
        g = ops.ext().multihot_probs_grad(idx, g_probs.float().contiguous(), ctx.num_experts, True)
        ops._count()
        return None, g.to(ctx.dtype), None


def fused_indices_to_multihot(indices: torch.Tensor, probs_indices: torch.Tensor, num_of_local_experts: int):
    """indices [T, k] (−1 = dropped), probs [T, k] → (multihot [T, E] bool, probs [T, E])."""
    if _use_kernels(indices, probs_indices):
        return _IndicesToMultihot.apply(indices, probs_indices, num_of_local_experts)
    T, k = indices.shape
    valid = indices >= 0
    rows = torch.arange(T, device=indices.device).unsqueeze(1).expand(T, k)[valid]
    cols = indices[valid].long()
    multihot = torch.zeros(T, num_of_local_experts, dtype=torch.bool, device=indices.device)
    multihot[rows, cols] = True
    probs = torch.zeros(T, num_of_local_experts, dtype=probs_indices.dtype, device=indices.device)
    probs = probs.index_put((rows, cols), probs_indices[valid])
    return multihot, probs


def fused_multihot_to_indices(multihot: torch.Tensor, probs: torch.Tensor, topk: int):
    """Inverse: routing map [T, E] → (indices [T, k] in expert order, padded with −1; probs [T, k])."""
    if _use_kernels(multihot, probs):
        return _MultihotToIndices.apply(multihot, probs, topk)
    E = multihot.shape[1]
    # stable: selected experts first, in expert order
    order = torch.argsort((~multihot.bool()).to(torch.int8), dim=1, stable=True)[:, :topk]
    if order.shape[1] < topk:
        order = torch.nn.functional.pad(order, (0, topk - order.shape[1]))
    picked = multihot.bool().gather(1, order) if E >= topk else torch.nn.functional.pad(multihot.bool().gather(1, order[:, :E]), (0, topk - E))
    idx = torch.where(picked, order, torch.full_like(order, -1))
    return idx, torch.where(picked, probs.gather(1, order.clamp(max=E - 1)), torch.zeros_like(probs[:, :1]).expand_as(idx))
