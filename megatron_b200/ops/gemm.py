"""GEMM front end: our tcgen05 kernels (4 tile/cluster variants) and cuBLAS, chosen by measurement.

``MEGATRON_B200_GEMM`` = ``auto`` (default: time every candidate once per distinct problem and keep the
fastest — "measure, don't guess"), ``tcgen05`` (our kernels only, heuristic variant) or ``cublas``.
Plain GEMMs may legitimately resolve to cuBLAS; the fused GEMM⇄collective kernels never do.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from . import reference as ref

_BACKEND = os.environ.get("MEGATRON_B200_GEMM", "auto")


def set_mode(mode: str) -> None:
    """Switch the plain-GEMM provider at run time (``auto`` | ``tcgen05[:variant]`` | ``cublas``); used by the batch-invariant mode."""
    global _BACKEND
    _BACKEND = mode
_CHOICE: Dict[Tuple, Tuple[str, int]] = {}
_TUNE_LOG = []
VARIANTS = {1: "1cta-128x256", 2: "1cta-128x128", 3: "2cta-256x256", 4: "2cta-256x128", 5: "2cta-256x256-tma-epilogue", 6: "2cta-256x128-tma-epilogue"}


def set_gemm_backend(name: str):
    global _BACKEND
    assert name in ("auto", "tcgen05", "cublas") or name.startswith("tcgen05:")
    _BACKEND = name


def get_gemm_backend() -> str:
    return _BACKEND


def tuning_report():
    """[(key, winner, {candidate: ms})] for every problem tuned so far."""
    return list(_TUNE_LOG)


def _ext():
    from . import ext

    return ext()


def _count():
    from . import _count as c

    c()


def _lib(layout: int, a, b, out, accumulate):
    if layout == 0:
        r = torch.matmul(a, b.t())
    elif layout == 1:
        r = torch.matmul(a, b)
    else:
        if accumulate and out is not None and out.dtype == a.dtype:
            return out.addmm_(a.t(), b)
        r = torch.matmul(a.t(), b)
    if out is None:
        return r
    if accumulate:
        out.add_(r)
    else:
        out.copy_(r)
    return out


def _ours(layout: int, a, b, out, accumulate, variant: int):
    _ext().gemm_bf16(a, b, out, layout, accumulate, variant)
    _count()
    return out


def _time(fn, iters=3) -> float:
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _tune(key, layout, a, b, out_shape, out_dtype, accumulate):
    scratch = torch.zeros(out_shape, dtype=out_dtype, device=a.device)
    res = {}
    for v in VARIANTS:
        try:
            res[f"tcgen05:{v}"] = _time(lambda: _ours(layout, a, b, scratch, accumulate, v))
        except RuntimeError:
            pass
    res["cublas"] = _time(lambda: _lib(layout, a, b, scratch, accumulate))
    best = min(res, key=res.get)
    choice = ("cublas", 0) if best == "cublas" else ("tcgen05", int(best.split(":")[1]))
    _CHOICE[key] = choice
    _TUNE_LOG.append((key, best, {k: round(v, 4) for k, v in res.items()}))
    return choice


def _dims(layout, a, b):
    if layout == 0:
        return a.shape[0], b.shape[0], a.shape[1]
    if layout == 1:
        return a.shape[0], b.shape[1], a.shape[1]
    return a.shape[1], b.shape[1], a.shape[0]


def _eligible(layout, a, b, out_dtype) -> bool:
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or out_dtype not in (torch.bfloat16, torch.float32):
        return False
    M, N, K = _dims(layout, a, b)
    # 16-byte row pitch for every TMA-loaded operand and for the vectorised epilogue stores
    return a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0 and N % 8 == 0 and M > 0 and N > 0 and K > 0


def gemm(layout: int, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False, out_dtype=None) -> torch.Tensor:
    """2-D GEMM in one of the three layouts (see ``csrc/gemm_sm100.cu``); ``a``/``b`` contiguous."""
    M, N, K = _dims(layout, a, b)
    od = out.dtype if out is not None else (out_dtype or a.dtype)
    if not a.is_cuda or _BACKEND == "cublas" or not _eligible(layout, a, b, od):
        if not a.is_cuda:
            return (ref.gemm_nt, ref.gemm_nn, None)[layout](a, b) if layout < 2 and out is None else _cpu(layout, a, b, out, accumulate, od)
        r = _lib(layout, a, b, out, accumulate)
        return r if r.dtype == od else r.to(od)
    if out is None:
        out = torch.empty((M, N), dtype=od, device=a.device)
        accumulate = False
    if _BACKEND.startswith("tcgen05"):
        v = int(_BACKEND.split(":")[1]) if ":" in _BACKEND else 0
        return _ours(layout, a, b, out, accumulate, v)
    key = (layout, M, N, K, bool(accumulate), od)
    choice = _CHOICE.get(key)
    if choice is None:
        choice = _tune(key, layout, a, b, (M, N), od, accumulate)
    if choice[0] == "cublas":
        return _lib(layout, a, b, out, accumulate)
    return _ours(layout, a, b, out, accumulate, choice[1])


def _cpu(layout, a, b, out, accumulate, od):
    if layout == 0:
        r = ref.gemm_nt(a, b)
    elif layout == 1:
        r = ref.gemm_nn(a, b)
    else:
        return ref.gemm_tn(a, b, out, accumulate, od)
    if out is not None:
        if accumulate:
            out.add_(r.to(out.dtype))
        else:
            out.copy_(r)
        return out
    return r.to(od)


def _flat(x):
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.is_contiguous() else x2.contiguous()


def gemm_nt(x: torch.Tensor, w: torch.Tensor, out_dtype=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``x[..., K] @ w[N, K]ᵀ`` → ``[..., N]`` (forward of every linear)."""
    x2 = _flat(x)
    o2 = out.view(x2.shape[0], w.shape[0]) if out is not None else None
    y = gemm(0, x2, w.contiguous(), o2, False, out_dtype)
    return y.view(*x.shape[:-1], w.shape[0])


def gemm_nn(gy: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``gy[..., N] @ w[N, K]`` → ``[..., K]`` (dgrad); ``out`` may be a symmetric-memory tensor."""
    g2 = _flat(gy)
    o2 = out.view(g2.shape[0], w.shape[1]) if out is not None else None
    y = gemm(1, g2, w.contiguous(), o2, False, None)
    return y.view(*gy.shape[:-1], w.shape[1])


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False, out_dtype=None) -> torch.Tensor:
    """``a[T, N]ᵀ @ b[T, K]`` → ``[N, K]`` (wgrad); ``accumulate`` adds into the fp32/bf16 ``out`` in the epilogue."""
    return gemm(2, a.contiguous(), b.contiguous(), out, accumulate and out is not None, out_dtype)
