"""Gradient norm / clipping / zero counting (reference ``optimizer/clip_grads.py:55-277``).

The norm is one multi-tensor sm_100a kernel; the clip coefficient stays on the device
and is folded into the fused Adam kernel's ``grad_scale`` — no host sync, no extra pass
over the gradients.
"""
from __future__ import annotations

from typing import List, Union

import torch
import torch.distributed as dist

from ... import ops
from ..utils import get_pg_size


def get_grad_norm_fp32(grads_for_norm: Union[List[torch.Tensor], torch.Tensor], norm_type: Union[int, float] = 2, grad_stats_parallel_group=None) -> torch.Tensor:
    """Global p-norm as a 0-d fp32 device tensor (reference returns a python float)."""
    if isinstance(grads_for_norm, torch.Tensor):
        grads_for_norm = [grads_for_norm]
    dev = grads_for_norm[0].device if grads_for_norm else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() != "gloo" else torch.device("cpu"))
    norm_type = float(norm_type)
    if norm_type == float("inf"):
        total = torch.stack([g.abs().max().float() for g in grads_for_norm]).max() if grads_for_norm else torch.zeros((), device=dev)
        total = total.reshape(1).to(dev)
        if grad_stats_parallel_group is not None and get_pg_size(grad_stats_parallel_group) > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX, group=grad_stats_parallel_group)
        return total[0]
    if norm_type == 2.0:
        n = ops.multi_tensor_l2norm(grads_for_norm).to(dev) if grads_for_norm else torch.zeros((), device=dev)
        total = (n.float() ** 2).reshape(1)
    else:
        total = torch.zeros(1, device=dev)
        for g in grads_for_norm:
            total += torch.linalg.vector_norm(g.float(), norm_type) ** norm_type
    if grad_stats_parallel_group is not None and get_pg_size(grad_stats_parallel_group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=grad_stats_parallel_group)
    return total[0] ** (1.0 / norm_type)


def clip_coefficient(total_norm: torch.Tensor, max_norm: float) -> torch.Tensor:
    """``min(1, max_norm / (norm + 1e-6))`` on device."""
    return torch.clamp(max_norm / (total_norm + 1.0e-6), max=1.0)


def clip_grad_by_total_norm_fp32(parameters, max_norm: Union[int, float], total_norm, use_decoupled_grad: bool = False):
    """In-place clip (API parity; the fused optimizers fold the coefficient into Adam instead)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = []
    for p in parameters:
        g = getattr(p, "decoupled_grad", None) if use_decoupled_grad else p.grad
        if g is not None:
            grads.append(g.detach())
    coeff = clip_coefficient(torch.as_tensor(total_norm, dtype=torch.float32, device=grads[0].device if grads else "cpu"), max_norm)
    ops.multi_tensor_scale(grads, coeff)


def count_zeros_fp32(parameters, grad_stats_parallel_group=None, use_decoupled_grad: bool = False, tp_group=None) -> torch.Tensor:
    from ..tensor_parallel import param_is_not_tensor_parallel_duplicate
    from ..transformer.module import param_is_not_shared

    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    total = None
    for p in parameters:
        g = getattr(p, "decoupled_grad", None) if use_decoupled_grad else p.grad
        if g is None or not param_is_not_shared(p) or not param_is_not_tensor_parallel_duplicate(p, tp_group):
            continue
        z = (g.numel() - torch.count_nonzero(g)).float()
        total = z if total is None else total + z
    if total is None:
        total = torch.zeros((), dtype=torch.float32)
    total = total.reshape(1)
    if grad_stats_parallel_group is not None and get_pg_size(grad_stats_parallel_group) > 1:
        if dist.get_backend(grad_stats_parallel_group) != "gloo" and not total.is_cuda:
            total = total.cuda()
        dist.all_reduce(total, group=grad_stats_parallel_group)
    return total[0]
