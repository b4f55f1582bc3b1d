"""Self check of every GEMM<->collective pair op against ``torch.distributed`` + an fp32 matmul.

``pair_op_self_check(group)`` runs each pair op of ``parallel/fused.py`` in the CURRENT execution mode (``fused`` on an
NVLink box) on Llama-3-8B-shaped operands and compares with the textbook formulation: NCCL/Gloo collective + fp32 GEMM.
Returned: ``{op: max|a-b| / max|b|, ..., "max": worst, "mode": resolved mode}``.  ``bench.py`` runs it before printing
its JSON line (``pair_op_max_rel_err``); ``tests/test_nvlink_gpu.py`` asserts on it, including back-to-back reuse of
the double-buffered workspaces and skewed ranks.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

from . import fused


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-20))


def _host_roundtrip(group, x):
    """Gloo control plane with CUDA data (several ranks time-slicing ONE GPU in the driver's 1-GPU test box): reference collectives run on the host."""
    return x.is_cuda and dist.get_backend(group) == "gloo"


def _ag(x, group):
    ws = dist.get_world_size(group)
    if _host_roundtrip(group, x):
        return _ag(x.cpu(), group).to(x.device)
    out = torch.empty((x.shape[0] * ws, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _rs(x, group):
    ws = dist.get_world_size(group)
    if _host_roundtrip(group, x):
        return _rs(x.cpu(), group).to(x.device)
    out = torch.empty((x.shape[0] // ws, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
    return out


def _ar(x, group):
    if _host_roundtrip(group, x):
        return _ar(x.cpu(), group).to(x.device)
    x = x.clone()
    dist.all_reduce(x, group=group)
    return x


def pair_op_self_check(group, seq: int = 8192, hidden: int = 4096, ffn: int = 14336, qkv: int = 6144, quick: bool = True, repeats: int = 1,
                       skew_rank: int = -1) -> Dict[str, float]:
    """See module docstring.  ``repeats`` > 1 re-runs every op back to back (workspace / flag reuse);
    ``skew_rank`` delays that rank before each op (late arrival of one peer)."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    from . import collectives

    on_gpu = torch.cuda.is_available() and (dist.get_backend(group) != "gloo" or collectives.backend_for(group) is not None)
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    g = torch.Generator(device=dev).manual_seed(4321 + rank)

    def rnd(*shape, scale=1.0):
        return (scale * torch.randn(*shape, device=dev, generator=g)).to(torch.bfloat16 if dev.type == "cuda" else torch.float32)

    def skew():
        if skew_rank == rank and dev.type == "cuda":
            torch.cuda._sleep(20_000_000)      # ~10 ms of device-side delay on one rank

    S = seq
    shapes = [("qkv", qkv // ws), ("fc1", 2 * ffn // ws)] if not quick else [("qkv", qkv // ws)]
    rshapes = [("proj", hidden // ws), ("fc2", ffn // ws)] if not quick else [("fc2", ffn // ws)]
    out: Dict[str, float] = {}

    def upd(name, val):
        out[name] = max(out.get(name, 0.0), val)

    for _ in range(repeats):
        for nm, n_local in shapes:                    # column-parallel layer: W [n_local, hidden]
            w = rnd(n_local, hidden, scale=0.02)
            xs = rnd(S // ws, 1, hidden)               # sequence-parallel shard
            gy = rnd(S, 1, n_local)
            xf = _ag(xs, group)
            skew()
            y = fused.all_gather_gemm(xs, w, group)
            upd(f"col_fwd_{nm}:ag_gemm", _rel(y, xf.float() @ w.float().t()))
            skew()
            gx, gw = fused.sp_linear_backward(gy, xs, w, group, True, False)
            upd(f"col_bwd_{nm}:gemm_rs", _rel(gx, _rs((gy.float() @ w.float()).to(gy.dtype).float(), group)))
            upd(f"col_bwd_{nm}:ag_wgrad", _rel(gw, gy.reshape(-1, n_local).float().t() @ xf.reshape(-1, hidden).float()))
            skew()
            gx2, h = fused.dgrad_all_reduce(gy, w, group)      # no-SP column dgrad
            if h is not None:
                h.wait()
            upd(f"col_bwd_{nm}:gemm_ar", _rel(gx2, _ar(gy.float() @ w.float(), group)))
        for nm, k_local in rshapes:                   # row-parallel layer: W [hidden, k_local]
            w = rnd(hidden, k_local, scale=0.02)
            x = rnd(S, 1, k_local)
            gys = rnd(S // ws, 1, hidden)
            ref = x.float() @ w.float().t()
            skew()
            y = fused.gemm_reduce_scatter(x, w, group)
            upd(f"row_fwd_{nm}:gemm_rs", _rel(y, _rs(ref, group)))
            skew()
            y2 = fused.gemm_all_reduce(x, w, group)
            upd(f"row_fwd_{nm}:gemm_ar", _rel(y2, _ar(ref, group)))
            skew()
            gx, gw = fused.row_linear_backward_sp(gys, x, w, group, True, False)
            gyf = _ag(gys, group)
            upd(f"row_bwd_{nm}:ag_gemm", _rel(gx, gyf.float() @ w.float()))
            upd(f"row_bwd_{nm}:ag_wgrad", _rel(gw, gyf.reshape(-1, hidden).float().t() @ x.reshape(-1, k_local).float()))
    worst = torch.tensor([max(out.values())], dtype=torch.float32, device="cpu" if dist.get_backend(group) == "gloo" else dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    res = {k: round(v, 6) for k, v in out.items()}
    res["max"] = round(float(worst), 6)
    res["mode"] = fused.get_mode(group)
    return res
