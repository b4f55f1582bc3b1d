"""Muon — momentum orthogonalised by Newton–Schulz (reference ``optimizer/muon.py`` + ``emerging_optimizers.py``, which wrap the external
``emerging_optimizers`` package; implemented here directly).

For a 2-D weight ``W [out, in]`` with momentum buffer ``M``:

    M ← β M + G;   U = G + β M (Nesterov) or M;   O = NewtonSchulz(U) ≈ U (UᵀU)^(-1/2);   W ← (1 - lr·wd) W - lr · scale · O

``NewtonSchulz`` runs the quintic iteration ``X ← a X + (b A + c A²) X`` with ``A = X Xᵀ`` on the bf16 tensor cores (5 steps, the published
coefficients), after normalising by the Frobenius norm so the singular values start inside the basin.  Everything that is not a hidden 2-D
weight — embeddings, output head, norms, biases, the MoE router — is updated with AdamW by the calling optimizer.

Tensor parallelism: ``blockwise`` orthogonalises every TP shard on its own (no communication; the update is block-orthogonal, which is what the
large-scale Muon runs use), ``duplicated`` all-gathers the momentum over the TP group, orthogonalises the full matrix redundantly and keeps the
local slice (exact, costs one all-gather per weight and step).  ZeRO-1 shards flat ranges, so Muon pairs with ``LayerWiseDistributedOptimizer``,
which hands WHOLE weights to data-parallel ranks.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

_NS_COEFFS = (3.4445, -4.7750, 2.0315)


def is_muon_param(param, master: Optional[torch.Tensor] = None) -> bool:
    t = master if master is not None else param
    if t.dim() != 2 or getattr(param, "is_embedding_or_output_parameter", False) or getattr(param, "is_router_parameter", False):
        return False
    name = getattr(param, "_muon_exclude", False)
    return not name


def newton_schulz(G: torch.Tensor, steps: int = 5, eps: float = 1e-7) -> torch.Tensor:
    """Approximate ``U Vᵀ`` of ``G = U S Vᵀ`` (the nearest semi-orthogonal matrix); accepts a leading batch of matrices."""
    a, b, c = _NS_COEFFS
    X = G.to(torch.bfloat16) if G.is_cuda else G.float()
    transpose = X.shape[-2] > X.shape[-1]
    if transpose:
        X = X.transpose(-1, -2)
    X = X / (X.norm(dim=(-2, -1), keepdim=True) + eps)
    for _ in range(steps):
        A = X @ X.transpose(-1, -2)
        B = b * A + c * (A @ A)
        X = a * X + B @ X
    if transpose:
        X = X.transpose(-1, -2)
    return X.to(G.dtype)


def _scale(shape, mode: str) -> float:
    out_f, in_f = shape[-2], shape[-1]
    if mode == "spectral":
        return max(1.0, out_f / in_f) ** 0.5
    if mode == "shape":
        return 0.2 * max(out_f, in_f) ** 0.5
    return 1.0


def orthogonalize(update: torch.Tensor, param=None, config=None) -> torch.Tensor:
    steps = getattr(config, "muon_ns_steps", 5)
    mode = getattr(config, "muon_tp_mode", "blockwise")
    tp = getattr(param, "tensor_model_parallel", False) if param is not None else False
    if tp and mode == "duplicated":
        from .. import parallel_state as ps

        group = ps.get_tensor_model_parallel_group()
        ws = dist.get_world_size(group)
        if ws > 1:
            dim = getattr(param, "partition_dim", 0)
            parts = [torch.empty_like(update) for _ in range(ws)]
            dist.all_gather(parts, update.contiguous(), group=group)
            full = newton_schulz(torch.cat(parts, dim=dim), steps)
            return full.chunk(ws, dim=dim)[dist.get_rank(group)].contiguous(), tuple(full.shape)
    return newton_schulz(update, steps), tuple(update.shape)


def muon_step(master: torch.Tensor, grad: torch.Tensor, momentum: torch.Tensor, lr: float, weight_decay: float, config=None, param=None) -> None:
    beta = getattr(config, "muon_momentum", 0.95)
    momentum.mul_(beta).add_(grad)
    upd = grad.add(momentum, alpha=beta) if getattr(config, "muon_nesterov", True) else momentum
    o, full_shape = orthogonalize(upd, param, config)
    s = _scale(full_shape, getattr(config, "muon_scale_mode", "spectral")) * getattr(config, "muon_extra_scale", 1.0)
    if weight_decay:
        master.mul_(1.0 - lr * weight_decay)
    master.add_(o.to(master.dtype), alpha=-lr * s)
