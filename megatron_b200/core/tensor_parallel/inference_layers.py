"""Inference-only tensor-parallel linears (reference ``tensor_parallel/inference_layers.py``): no autograd bookkeeping, weights used as stored, and the
decode-time collective fused with what follows it.

``InferenceRowParallelLinear.forward(x, residual, norm)`` covers the hot sequence of a decode step under TP + SP:

    partial = x · Wᵀ  →  reduce-scatter  →  h = residual + y  →  RMSNorm(h)  →  all-gather (input of the next column-parallel GEMM)

On CUDA with the NVLink backend the reduce-scatter / all-gather are the multimem kernels of ``parallel/nvlink.py`` and the middle is ONE kernel
(``ops.add_rms_norm``); on CPU / NCCL-only setups the same function runs through ``torch.distributed``.  Small decode batches are latency-bound
(≈ 10 µs per collective), so everything stays on one stream and is CUDA-graph capturable (no host syncs, static shapes)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..utils import get_pg_size
from .layers import ColumnParallelLinear, RowParallelLinear


def _all_gather_first_dim(x: torch.Tensor, group) -> torch.Tensor:
    ws = get_pg_size(group)
    if ws == 1:
        return x
    out = x.new_empty((x.shape[0] * ws,) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _reduce_scatter_first_dim(x: torch.Tensor, group) -> torch.Tensor:
    ws = get_pg_size(group)
    if ws == 1:
        return x
    out = x.new_empty((x.shape[0] // ws,) + tuple(x.shape[1:]))
    dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
    return out


class InferenceColumnParallelLinear(ColumnParallelLinear):
    @torch.no_grad()
    def forward(self, input_, weight=None, runtime_gather_output=None, input_is_gathered: bool = False):
        w = self.weight if weight is None else weight
        x = input_ if (input_is_gathered or not self.sequence_parallel) else _all_gather_first_dim(input_, self.tp_group)
        from ... import ops

        out = ops.gemm_nt(x, w) if x.is_cuda else torch.matmul(x, w.t())
        if self.bias is not None and not self.skip_bias_add:
            out = out + self.bias
        gather = self.gather_output if runtime_gather_output is None else runtime_gather_output
        if gather and get_pg_size(self.tp_group) > 1:
            parts = [torch.empty_like(out) for _ in range(get_pg_size(self.tp_group))]
            dist.all_gather(parts, out.contiguous(), group=self.tp_group)
            out = torch.cat(parts, dim=-1)
        return out, (self.bias if self.skip_bias_add else None)


class InferenceRowParallelLinear(RowParallelLinear):
    @torch.no_grad()
    def forward(self, input_, residual: Optional[torch.Tensor] = None, norm: Optional[torch.nn.Module] = None, gather_output: bool = True):
        """Without ``residual``: plain row-parallel linear (all-reduce, or reduce-scatter under SP) → ``(out, bias)``.
        With ``residual`` (sequence-sharded ``[s/tp, b, h]``) and ``norm`` (RMSNorm module): returns ``(normed, new_residual)`` where ``normed`` is the
        all-gathered ``[s, b, h]`` input of the next layer and ``new_residual`` stays sequence-sharded."""
        from ... import ops

        x = input_
        part = ops.gemm_nt(x, self.weight) if x.is_cuda else torch.matmul(x, self.weight.t())
        ws = get_pg_size(self.tp_group)
        if residual is None:
            if ws > 1:
                if self.sequence_parallel:
                    part = _reduce_scatter_first_dim(part, self.tp_group)
                else:
                    dist.all_reduce(part, group=self.tp_group)
            if not self.skip_bias_add and self.bias is not None:
                part = part + self.bias
            return part, (self.bias if self.skip_bias_add else None)
        y = _reduce_scatter_first_dim(part, self.tp_group)
        if self.bias is not None:
            y = y + self.bias
        w, eps, zc = norm.weight, norm.eps, getattr(norm, "zero_centered_gamma", False)
        normed, h = ops.add_rms_norm(y, residual, w, eps, zc)
        return (_all_gather_first_dim(normed, self.tp_group) if gather_output else normed), h


def convert_to_inference_layers(model: torch.nn.Module) -> torch.nn.Module:
    """Re-class the TP linears of a trained model in place (parameters are shared, nothing is copied)."""
    for m in model.modules():
        if type(m) is ColumnParallelLinear:
            m.__class__ = InferenceColumnParallelLinear
        elif type(m) is RowParallelLinear:
            m.__class__ = InferenceRowParallelLinear
    return model
