"""Autograd-aware collectives between tensor-parallel regions.

Capability parity with reference ``tensor_parallel/mappings.py`` (12 public
mapping functions, ``_AllToAll`` :424, ``all_to_all_sp2hp/hp2sp`` :566-621).
Every mapping is declared as a (forward primitive, backward primitive) pair in
one table and materialised into an autograd Function by ``_make_mapping`` —
the conjugate pairs (copy↔reduce, scatter↔gather, all-gather↔reduce-scatter)
are explicit instead of being spread over nine hand-written classes.

On GPU ranks with a symmetric heap the primitives route to the NVLink kernels
in ``megatron_b200.parallel.collectives`` (multimem all-gather / reduce-scatter
/ all-reduce); otherwise they use ``torch.distributed`` (Gloo on CPU).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from ..utils import get_pg_rank, get_pg_size, get_tensor_model_parallel_group_if_none

# -----------------------------------------------------------------------------
# primitives (no autograd)
# -----------------------------------------------------------------------------


def _size(group) -> int:
    return get_pg_size(group) if group is not None else 1


def _nvlink(group, t: torch.Tensor):
    """Return the NVLink collective backend for ``group`` if one is active."""
    if not t.is_cuda:
        return None
    from ...parallel import collectives as nvl

    return nvl.backend_for(group)


def _reduce(x, group):
    if _size(group) == 1:
        return x
    be = _nvlink(group, x)
    if be is not None:
        return be.all_reduce(x)
    x = x.contiguous()
    dist.all_reduce(x, group=group)
    return x


def _split_along_last_dim(x, group):
    ws = _size(group)
    if ws == 1:
        return x
    chunk = x.shape[-1] // ws
    r = get_pg_rank(group)
    return x[..., r * chunk : (r + 1) * chunk].contiguous()


def _split_along_first_dim(x, group):
    ws = _size(group)
    if ws == 1:
        return x
    assert x.shape[0] % ws == 0, "first dimension must be divisible by the group size"
    chunk = x.shape[0] // ws
    r = get_pg_rank(group)
    return x[r * chunk : (r + 1) * chunk].contiguous()


def _gather_along_last_dim(x, group):
    ws = _size(group)
    if ws == 1:
        return x
    full = _gather_along_first_dim(x.contiguous(), group)  # [ws*d0, ..., dl]
    parts = full.view(ws, *x.shape)
    return torch.cat(list(parts.unbind(0)), dim=-1).contiguous()


def _reduce_scatter_along_last_dim(x, group):
    ws = _size(group)
    if ws == 1:
        return x
    parts = torch.stack(x.chunk(ws, dim=-1), dim=0).contiguous()  # [ws, ..., dl/ws]
    flat = parts.view(ws * parts.shape[1], *parts.shape[2:]) if parts.dim() > 2 else parts.view(-1)
    out = _reduce_scatter_along_first_dim(flat, group)
    return out.view(*x.shape[:-1], x.shape[-1] // ws)


def _gather_along_first_dim(x, group, output_split_sizes: Optional[List[int]] = None, use_global_buffer: bool = False):
    ws = _size(group)
    if ws == 1:
        return x
    x = x.contiguous()
    if output_split_sizes is not None:
        out = torch.empty((sum(output_split_sizes), *x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather(list(torch.split(out, output_split_sizes, dim=0)), x, group=group)
        return out
    be = _nvlink(group, x)
    if be is not None:
        return be.all_gather(x)
    shape = (x.shape[0] * ws, *x.shape[1:])
    if use_global_buffer:
        from .. import parallel_state as ps

        out = ps.get_global_memory_buffer().get_tensor(shape, x.dtype, "mpu", device=x.device)
    else:
        out = torch.empty(shape, dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def _reduce_scatter_along_first_dim(x, group, input_split_sizes: Optional[List[int]] = None, use_global_buffer: bool = False):
    ws = _size(group)
    if ws == 1:
        return x
    x = x.contiguous()
    if input_split_sizes is not None:
        r = get_pg_rank(group)
        out = torch.empty((input_split_sizes[r], *x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter(out, list(torch.split(x, input_split_sizes, dim=0)), group=group)
        return out
    assert x.shape[0] % ws == 0, "first dimension must be divisible by the group size"
    be = _nvlink(group, x)
    if be is not None:
        return be.reduce_scatter(x)
    out = torch.empty((x.shape[0] // ws, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x, group=group)
    return out


def _identity(x, group):
    return x


# -----------------------------------------------------------------------------
# autograd wrappers generated from (fwd, bwd) primitive pairs
# -----------------------------------------------------------------------------


def _make_mapping(name: str, fwd: Callable, bwd: Callable):
    class _Mapping(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, group, *extra):
            ctx.group, ctx.extra = group, extra
            return fwd(x, group, *extra)

        @staticmethod
        def backward(ctx, g):
            return (bwd(g, ctx.group, *ctx.extra), None) + (None,) * len(ctx.extra)

    _Mapping.__name__ = _Mapping.__qualname__ = name
    return _Mapping


_CopyToModelParallelRegion = _make_mapping("_CopyToModelParallelRegion", _identity, _reduce)
_ReduceFromModelParallelRegion = _make_mapping("_ReduceFromModelParallelRegion", _reduce, _identity)
_ScatterToModelParallelRegion = _make_mapping("_ScatterToModelParallelRegion", _split_along_last_dim, _gather_along_last_dim)
_GatherFromModelParallelRegion = _make_mapping("_GatherFromModelParallelRegion", _gather_along_last_dim, _split_along_last_dim)
_ScatterToSequenceParallelRegion = _make_mapping("_ScatterToSequenceParallelRegion", _split_along_first_dim, _gather_along_first_dim)
_ReduceScatterToSequenceParallelRegion = _make_mapping(
    "_ReduceScatterToSequenceParallelRegion", _reduce_scatter_along_first_dim, _gather_along_first_dim
)
_AllGatherFromTensorParallelRegion = _make_mapping(
    "_AllGatherFromTensorParallelRegion", _gather_along_last_dim, _reduce_scatter_along_last_dim
)
_ReduceScatterToTensorParallelRegion = _make_mapping(
    "_ReduceScatterToTensorParallelRegion", _reduce_scatter_along_last_dim, _gather_along_last_dim
)


class _GatherFromSequenceParallelRegion(torch.autograd.Function):
    """all-gather along dim 0; backward is reduce-scatter (TP-shared consumer) or split."""

    @staticmethod
    def forward(ctx, x, group, tensor_parallel_output_grad, output_split_sizes, use_global_buffer):
        ctx.group = group
        ctx.rs_in_bwd = tensor_parallel_output_grad
        ctx.splits = output_split_sizes
        ctx.use_global_buffer = use_global_buffer
        return _gather_along_first_dim(x, group, output_split_sizes, use_global_buffer)

    @staticmethod
    def backward(ctx, g):
        if ctx.rs_in_bwd:
            out = _reduce_scatter_along_first_dim(g, ctx.group, ctx.splits, ctx.use_global_buffer)
        else:
            assert ctx.splits is None
            out = _split_along_first_dim(g, ctx.group)
        return out, None, None, None, None


class _ReduceScatterToSequenceParallelRegionV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, input_split_sizes, use_global_buffer):
        ctx.group, ctx.splits, ctx.use_global_buffer = group, input_split_sizes, use_global_buffer
        return _reduce_scatter_along_first_dim(x, group, input_split_sizes, use_global_buffer)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g, ctx.group, ctx.splits, ctx.use_global_buffer), None, None, None


class _AllToAll(torch.autograd.Function):
    """Variable-size all-to-all along dim 0 (reference :424-489)."""

    @staticmethod
    def forward(ctx, group, x, output_split_sizes, input_split_sizes):
        ctx.group, ctx.out_splits, ctx.in_splits = group, output_split_sizes, input_split_sizes
        ws = _size(group)
        if ws == 1:
            return x
        x = x.contiguous()
        if output_split_sizes is None:
            out = torch.empty_like(x)
        else:
            out = x.new_empty((sum(output_split_sizes), *x.shape[1:]))
        if x.is_cuda or dist.get_backend(group) != "gloo":
            dist.all_to_all_single(out, x, output_split_sizes=output_split_sizes, input_split_sizes=input_split_sizes, group=group)
        else:
            # gloo has no all_to_all_single for uneven splits on every build: emulate with all_to_all
            ins = list(x.split(input_split_sizes if input_split_sizes is not None else x.shape[0] // ws, dim=0))
            outs = list(out.split(output_split_sizes if output_split_sizes is not None else out.shape[0] // ws, dim=0))
            ins = [t.contiguous() for t in ins]
            outs_c = [torch.empty_like(t) for t in outs]
            ranks = dist.get_process_group_ranks(group)
            me = dist.get_rank(group)
            outs_c[me].copy_(ins[me])
            reqs = []
            for peer in range(ws):
                if peer != me:
                    reqs.append(dist.irecv(outs_c[peer], ranks[peer], group=group))
            for peer in range(ws):
                if peer != me:
                    reqs.append(dist.isend(ins[peer], ranks[peer], group=group))
            for r in reqs:
                r.wait()
            for o, oc in zip(outs, outs_c):
                o.copy_(oc)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, _AllToAll.apply(ctx.group, g, ctx.in_splits, ctx.out_splits), None, None


# -----------------------------------------------------------------------------
# public functional API (names match the reference)
# -----------------------------------------------------------------------------


def _tp(group):
    return get_tensor_model_parallel_group_if_none(group)


def copy_to_tensor_model_parallel_region(input_, group=None):
    return _CopyToModelParallelRegion.apply(input_, _tp(group))


def reduce_from_tensor_model_parallel_region(input_, group=None):
    return _ReduceFromModelParallelRegion.apply(input_, _tp(group))


def scatter_to_tensor_model_parallel_region(input_, group=None):
    return _ScatterToModelParallelRegion.apply(input_, _tp(group))


def gather_from_tensor_model_parallel_region(input_, group=None):
    return _GatherFromModelParallelRegion.apply(input_, _tp(group))


def scatter_to_sequence_parallel_region(input_, group=None):
    return _ScatterToSequenceParallelRegion.apply(input_, _tp(group))


def gather_from_sequence_parallel_region(
    input_, tensor_parallel_output_grad=True, group=None, output_split_sizes=None, use_global_buffer=False
):
    return _GatherFromSequenceParallelRegion.apply(
        input_, _tp(group), tensor_parallel_output_grad, output_split_sizes, use_global_buffer
    )


def reduce_scatter_to_sequence_parallel_region(input_, group=None, input_split_sizes=None, use_global_buffer=False):
    if input_split_sizes is None and not use_global_buffer:
        return _ReduceScatterToSequenceParallelRegion.apply(input_, _tp(group))
    return _ReduceScatterToSequenceParallelRegionV.apply(input_, _tp(group), input_split_sizes, use_global_buffer)


def all_gather_last_dim_from_tensor_parallel_region(input_, group=None):
    return _AllGatherFromTensorParallelRegion.apply(input_, _tp(group))


def reduce_scatter_last_dim_to_tensor_parallel_region(input_, group=None):
    return _ReduceScatterToTensorParallelRegion.apply(input_, _tp(group))


def all_to_all(group, input_, output_split_sizes_=None, input_split_sizes=None):
    return _AllToAll.apply(group, input_, output_split_sizes_, input_split_sizes)


def all_to_all_sp2hp(input_, group=None):
    """[s/tp, H] → [s, H/tp]: swap sequence sharding for hidden sharding."""
    group = _tp(group)
    ws = _size(group)
    if ws == 1:
        return input_
    x = input_.reshape(-1, input_.shape[-1])
    parts = torch.cat(torch.split(x, x.shape[-1] // ws, dim=1), dim=0)  # [ws * s/tp, H/tp]
    return all_to_all(group, parts)


def all_to_all_hp2sp(input_, group=None):
    """[s, H/tp] → [s/tp, H]."""
    group = _tp(group)
    ws = _size(group)
    if ws == 1:
        return input_
    x = input_.reshape(-1, input_.shape[-1])
    y = all_to_all(group, x)
    return torch.cat(torch.split(y, y.shape[0] // ws, dim=0), dim=1)
