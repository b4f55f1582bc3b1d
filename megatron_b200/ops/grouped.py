"""Grouped GEMM over per-expert token groups.

Rows of ``x`` are grouped by expert (``tokens_per_expert[e]`` consecutive rows for expert e);
weights are stacked ``w[e]``.  On B200 the native grouped kernel (``csrc/grouped_gemm_sm100.cu``)
walks all (expert, tile) pairs in one persistent launch; otherwise each group goes through the
dense tcgen05 GEMM / torch.matmul.  Replaces TE ``GroupedLinear``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _use_cuda, ext, gemm_nn, gemm_nt, gemm_tn, _count


def _offsets(tpe: Sequence[int]) -> List[int]:
    out, acc = [0], 0
    for n in tpe:
        acc += int(n)
        out.append(acc)
    return out


def _native(x):
    return _use_cuda(x) and hasattr(ext(), "grouped_gemm_bf16") and x.dtype == torch.bfloat16


def grouped_gemm_nt(x: torch.Tensor, w: torch.Tensor, tokens_per_expert: Sequence[int]) -> torch.Tensor:
    """``out[g_e] = x[g_e] @ w[e]ᵀ``;  x [T, K], w [E, N, K] → [T, N]."""
    off = _offsets(tokens_per_expert)
    out = torch.empty((x.shape[0], w.shape[1]), dtype=x.dtype, device=x.device)
    if _native(x) and x.shape[1] % 8 == 0 and w.shape[1] % 8 == 0:
        ext().grouped_gemm_bf16(x.contiguous(), w.contiguous(), out, torch.tensor(off, dtype=torch.int32), 0, False)
        _count()
        return out
    for e in range(w.shape[0]):
        if off[e + 1] > off[e]:
            out[off[e] : off[e + 1]] = gemm_nt(x[off[e] : off[e + 1]], w[e])
    return out


def grouped_gemm_nn(gy: torch.Tensor, w: torch.Tensor, tokens_per_expert: Sequence[int]) -> torch.Tensor:
    """dgrad: ``gx[g_e] = gy[g_e] @ w[e]``;  gy [T, N], w [E, N, K] → [T, K]."""
    off = _offsets(tokens_per_expert)
    out = torch.empty((gy.shape[0], w.shape[2]), dtype=gy.dtype, device=gy.device)
    if _native(gy) and gy.shape[1] % 8 == 0 and w.shape[2] % 8 == 0:
        ext().grouped_gemm_bf16(gy.contiguous(), w.contiguous(), out, torch.tensor(off, dtype=torch.int32), 1, False)
        _count()
        return out
    for e in range(w.shape[0]):
        if off[e + 1] > off[e]:
            out[off[e] : off[e + 1]] = gemm_nn(gy[off[e] : off[e + 1]], w[e])
    return out


def grouped_gemm_tn(gy: torch.Tensor, x: torch.Tensor, tokens_per_expert: Sequence[int], w_like: torch.Tensor) -> torch.Tensor:
    """wgrad: ``gw[e] = gy[g_e]ᵀ @ x[g_e]`` → [E, N, K]."""
    off = _offsets(tokens_per_expert)
    gw = torch.zeros_like(w_like)
    if _native(gy) and gy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0:
        ext().grouped_gemm_bf16(gy.contiguous(), x.contiguous(), gw, torch.tensor(off, dtype=torch.int32), 2, False)
        _count()
        return gw
    for e in range(w_like.shape[0]):
        if off[e + 1] > off[e]:
            gw[e] = gemm_tn(gy[off[e] : off[e + 1]], x[off[e] : off[e + 1]], out_dtype=w_like.dtype)
    return gw
