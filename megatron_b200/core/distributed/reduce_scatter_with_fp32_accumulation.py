"""Reduce-scatter of low-precision gradients with fp32 accumulation (reference ``distributed/reduce_scatter_with_fp32_accumulation.py``).

A bf16 reduce-scatter over ``n`` ranks adds ``n`` bf16 numbers pairwise in bf16 and loses ≈ log2(n) bits; here every rank receives the ``n`` bf16
shards that belong to it with ONE all-to-all and sums them in fp32 locally — the wire format stays 2 bytes/element, only the arithmetic is wider.
(The NVLink path gets the same effect from ``multimem.ld_reduce … .acc::f32`` inside the switch, ``ops/csrc/nvlink_collectives.cu``.)"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def reduce_scatter_with_fp32_accumulation(output: torch.Tensor, input: torch.Tensor, group=None, scale: Optional[float] = None, async_op: bool = False):
    """``output [n/ws]`` (any float dtype, typically fp32) ← Σ_ranks ``input [n]`` (bf16/fp16) shard of this rank, accumulated in fp32."""
    ws = dist.get_world_size(group)
    assert input.numel() % ws == 0 and output.numel() == input.numel() // ws
    if ws == 1:
        output.copy_(input.float() if scale is None else input.float() * scale)
        return None
    recv = torch.empty_like(input)
    work = dist.all_to_all_single(recv, input.contiguous(), group=group, async_op=async_op)

    def finish():
        acc = recv.view(ws, -1).float().sum(dim=0)
        if scale is not None:
            acc.mul_(scale)
        output.copy_(acc.view_as(output))

    if not async_op:
        finish()
        return None

    class _Handle:
        def wait(self_inner):
            work.wait()
            finish()

    return _Handle()
