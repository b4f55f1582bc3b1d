"""NVFP4 (E2M1 values, one E4M3 scale per 16-element block + one fp32 tensor scale) quantisation helpers (reference ``core/fp4_utils.py``).

``quantize_nvfp4`` / ``dequantize_nvfp4`` define the numerics; ``nvfp4_linear`` runs W4A4 on the block-scaled tensor-core GEMM
(``ops.gemm_nvfp4_nt`` → ``csrc/gemm_nvfp4_sm100.cu``: ``tcgen05.mma kind::mxf4nvf4.block_scale.block16``, scale factors in TMEM) when the shapes allow
(CUDA, K % 256 == 0) with the activation quantised on the fly, or weight-only (dequantise + bf16 GEMM) otherwise."""
from __future__ import annotations

from typing import Tuple

import torch

_E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
BLOCK = 16


def is_nvfp4tensor(t) -> bool:
    return isinstance(t, tuple) and len(t) == 3 and getattr(t[0], "dtype", None) == torch.uint8


def quantize_nvfp4(x: torch.Tensor, stochastic: bool = False, generator=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x [..., K] (K % 16 == 0) → (codes uint8 [..., K] sign<<3|index, block scales fp8-e4m3 [..., K/16], tensor scale fp32).
    ``stochastic``: round each magnitude up or down to the neighbouring grid points with probability proportional to proximity — unbiased
    (E[q] = x), which is what gradients need at 4 bits (nearest rounding systematically loses the small components)."""
    assert x.shape[-1] % BLOCK == 0
    xf = x.float()
    tscale = xf.abs().amax().clamp(min=1e-12) / (6.0 * 448.0)
    xb = (xf / tscale).view(*x.shape[:-1], -1, BLOCK)
    bscale = (xb.abs().amax(-1, keepdim=True) / 6.0).clamp(min=2.0**-9).to(torch.float8_e4m3fn)
    y = xb / bscale.float()
    grid = _E2M1.to(x.device)
    if stochastic:
        a = y.abs().clamp(max=6.0)
        hi = torch.searchsorted(grid, a.contiguous(), right=False).clamp(1, 7)         # first grid point >= a
        lo = hi - 1
        g_lo, g_hi = grid[lo], grid[hi]
        p_hi = ((a - g_lo) / (g_hi - g_lo)).clamp(0, 1)
        u = torch.rand(a.shape, device=a.device, generator=generator)
        idx = torch.where(u < p_hi, hi, lo)
    else:
        idx = (y.abs().unsqueeze(-1) - grid).abs().argmin(-1)
    codes = (idx | ((y < 0).to(torch.int64) << 3)).to(torch.uint8)
    return codes.view(x.shape), bscale.squeeze(-1), tscale


def dequantize_nvfp4(codes: torch.Tensor, bscale: torch.Tensor, tscale: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    grid = _E2M1.to(codes.device)
    mag = grid[(codes & 7).long()]
    val = torch.where((codes & 8) > 0, -mag, mag).view(*codes.shape[:-1], -1, BLOCK)
    return (val * bscale.float().unsqueeze(-1) * tscale).view(codes.shape).to(dtype)


def nvfp4_linear(x: torch.Tensor, qweight, quantize_activation: bool = True) -> torch.Tensor:
    """``x @ Wᵀ`` with a weight stored as NVFP4 ``(codes, block scales, tensor scale)``.  ``quantize_activation`` → W4A4 on the fp4 tensor cores."""
    K = x.shape[-1]
    if quantize_activation and x.is_cuda and K % 256 == 0:
        from .. import ops

        x2 = x.reshape(-1, K)
        out = ops.gemm_nvfp4_nt(*ops.nvfp4_quantize(x2), *qweight)
        return out.to(x.dtype).view(*x.shape[:-1], out.shape[-1])
    return torch.nn.functional.linear(x, dequantize_nvfp4(*qweight, dtype=x.dtype))


def hadamard16(x: torch.Tensor) -> torch.Tensor:
    """Orthonormal 16-point Walsh-Hadamard transform on consecutive groups of 16 along the last dim (its own inverse).  Applied to BOTH operands of a
    GEMM along the reduction dimension it leaves the product unchanged (Hᵀ H = I) while spreading outliers over the 16-element scaling block."""
    shp = x.shape
    y = x.float().reshape(-1, 16)
    h = 1
    while h < 16:
        y = y.view(-1, 16 // (2 * h), 2, h)
        y = torch.stack([y[:, :, 0] + y[:, :, 1], y[:, :, 0] - y[:, :, 1]], dim=2).reshape(-1, 16)
        h *= 2
    return (y * 0.25).view(shp).to(x.dtype)
