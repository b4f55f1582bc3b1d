"""Pair ops: a GEMM together with the collective adjacent to it.

This is the seam between the model code (``core/tensor_parallel/layers.py``)
and the three execution modes:

``"gloo"/"nccl"``  collective via ``torch.distributed`` + GEMM (CPU tests, baseline mode)
``"nvlink"``       our NVLink multimem/P2P collective kernels on a side stream + tcgen05 GEMM
``"fused"``        ONE sm_100a kernel per pair op: comm CTAs move/reduce tiles over
                    NVSwitch while tcgen05 CTAs compute (``ops/csrc/fused_tp_gemm.cu``)

The reference expresses the same data flow as separate NCCL calls around
``torch.matmul`` (``tensor_parallel/layers.py:623-631, 666-712, 1561-1564``).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops
from ..core.utils import get_pg_size

_MODE = os.environ.get("MEGATRON_B200_TP_COMM", "auto")  # auto | nccl | nvlink | fused
# what "auto" means on this build: the fastest MEASURED mode per TP size on B200 + NVSwitch (profiles/r1_tp_comm.md):
# fused vs NCCL tokens/s on Llama-3 8B:  TP=8 107.3k vs 82.5k (1.30x);  TP=4 65.3k vs 59.4k (1.10x);  TP=2 36.5k vs 35.2k (1.03x).
_AUTO_FUSED_MIN_TP = int(os.environ.get("MEGATRON_B200_FUSED_MIN_TP", "2"))


def set_mode(mode: str) -> None:
    global _MODE
    assert mode in ("auto", "nccl", "nvlink", "fused")
    _MODE = mode


def get_mode(group=None, world_size: Optional[int] = None) -> str:
    """Resolved TP-communication mode for a group (or an explicit TP size)."""
    if _MODE != "auto":
        return _MODE
    if world_size is None:
        if group is not None:
            world_size = get_pg_size(group)
        else:
            from ..core import parallel_state as ps

            world_size = ps.get_tensor_model_parallel_world_size() if ps.model_parallel_is_initialized() else 1
    return "fused" if world_size >= _AUTO_FUSED_MIN_TP else "nccl"


def _nvl(group, t):
    """NVLink backend (symmetric-heap collectives) or None."""
    if not t.is_cuda or get_mode(group) == "nccl":
        return None
    from . import collectives

    return collectives.backend_for(group)


# ---- plain GEMMs ---------------------------------------------------------------


def gemm_nt(x: torch.Tensor, w: torch.Tensor, out_dtype=None) -> torch.Tensor:
    """``x[..., K] @ w[N, K]ᵀ``."""
    return ops.gemm_nt(x, w, out_dtype=out_dtype)


def gemm_nn(gy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``gy[..., N] @ w[N, K]`` (dgrad)."""
    return ops.gemm_nn(gy, w)


_DUMMY_WGRAD = {}


def _dummy_grad(weight: torch.Tensor) -> torch.Tensor:
    key = (tuple(weight.shape), weight.dtype, weight.device)
    d = _DUMMY_WGRAD.get(key)
    if d is None:
        d = torch.empty(1, dtype=weight.dtype, device=weight.device).expand(weight.shape)
        _DUMMY_WGRAD[key] = d
    return d


def wgrad(gy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, accumulate_into_main_grad: bool) -> Optional[torch.Tensor]:
    """``gyᵀ @ x`` → [N, K]; optionally accumulated straight into fp32 ``weight.main_grad``
    by the GEMM epilogue (gradient-accumulation fusion, reference X2)."""
    gy2 = gy.reshape(-1, gy.shape[-1])
    x2 = x.reshape(-1, x.shape[-1])
    if accumulate_into_main_grad and hasattr(weight, "main_grad"):
        ops.gemm_tn(gy2, x2, out=weight.main_grad, accumulate=True)
        weight.grad_added_to_main_grad = True
        return _dummy_grad(weight)
    return ops.gemm_tn(gy2, x2, out_dtype=weight.dtype)


# ---- collectives wrapped so both back ends look alike ----------------------------


class _Done:
    def wait(self):
        return None


def all_reduce_async(t: torch.Tensor, group):
    be = _nvl(group, t)
    if be is not None:
        return be.all_reduce_async(t)
    return dist.all_reduce(t, group=group, async_op=True)


def _all_gather_first(x: torch.Tensor, group, async_op=False):
    ws = get_pg_size(group)
    out = torch.empty((x.shape[0] * ws, *x.shape[1:]), dtype=x.dtype, device=x.device)
    h = dist.all_gather_into_tensor(out, x.contiguous(), group=group, async_op=async_op)
    return out, h


def _reduce_scatter_first(x: torch.Tensor, group, async_op=False):
    ws = get_pg_size(group)
    out = torch.empty((x.shape[0] // ws, *x.shape[1:]), dtype=x.dtype, device=x.device)
    h = dist.reduce_scatter_tensor(out, x.contiguous(), group=group, async_op=async_op)
    return out, h


# ---- pair ops ------------------------------------------------------------------


def all_gather_gemm(x: torch.Tensor, w: torch.Tensor, group) -> torch.Tensor:
    """Sequence-parallel column linear forward: AG(x over dim 0) then ``X Wᵀ``."""
    be = _nvl(group, x)
    if be is not None:
        return be.all_gather_gemm(x, w)
    full, _ = _all_gather_first(x, group)
    return gemm_nt(full, w)


def gemm_reduce_scatter(x: torch.Tensor, w: torch.Tensor, group) -> torch.Tensor:
    """Row linear forward under SP: ``X Wᵀ`` then RS over dim 0."""
    be = _nvl(group, x)
    if be is not None:
        return be.gemm_reduce_scatter(x, w)
    y = gemm_nt(x, w)
    out, _ = _reduce_scatter_first(y, group)
    return out


def gemm_all_reduce(x: torch.Tensor, w: torch.Tensor, group) -> torch.Tensor:
    """Row linear forward without SP: ``X Wᵀ`` then all-reduce (one fused kernel on the NVLink backend)."""
    be = _nvl(group, x)
    if be is not None:
        return be.gemm_all_reduce(x, w, 0)
    y = gemm_nt(x, w)
    dist.all_reduce(y, group=group)
    return y


def dgrad_all_reduce(gy: torch.Tensor, w: torch.Tensor, group):
    """Column linear dgrad without SP: ``dY W`` then all-reduce.  Returns ``(gx, handle)``: the fused kernel has already
    reduced when it returns (handle None); the NCCL path returns the async handle so the wgrad GEMM overlaps the collective."""
    be = _nvl(group, gy)
    if be is not None:
        return be.gemm_all_reduce(gy, w, 1), None
    gx = gemm_nn(gy, w)
    return gx, dist.all_reduce(gx, group=group, async_op=True)


def sp_linear_backward(gy, x, weight, group, wgrad_needed: bool, accumulate: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Backward of the SP column linear.

    dgrad:  ``gy @ W`` → reduce-scatter      (GEMM→RS pair)
    wgrad:  all-gather(x) → ``gyᵀ @ X``      (AG→GEMM pair, fp32 accumulate epilogue)
    The two pairs are independent, so the AG of ``x`` hides under the dgrad GEMM
    and the RS of dgrad hides under the wgrad GEMM.
    """
    be = _nvl(group, x)
    if be is not None:
        return be.sp_linear_backward(gy, x, weight, wgrad_needed, accumulate, wgrad)
    full_x, h_ag = (None, None)
    if wgrad_needed:
        full_x, h_ag = _all_gather_first(x, group, async_op=True)
    gx_full = gemm_nn(gy, weight)
    gx, h_rs = _reduce_scatter_first(gx_full, group, async_op=True)
    gw = None
    if wgrad_needed:
        h_ag.wait()
        gw = wgrad(gy, full_x, weight, accumulate)
    h_rs.wait()
    return gx, gw


def row_linear_backward_sp(gy, x, weight, group, wgrad_needed: bool, accumulate: bool):
    """Backward of the SP row linear: all-gather(dY) feeds dgrad and wgrad."""
    be = _nvl(group, x)
    if be is not None:
        return be.row_linear_backward_sp(gy, x, weight, wgrad_needed, accumulate, wgrad)
    full_gy, _ = _all_gather_first(gy, group)
    gx = gemm_nn(full_gy, weight)
    gw = wgrad(full_gy, x, weight, accumulate) if wgrad_needed else None
    return gx, gw
