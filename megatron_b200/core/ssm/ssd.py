"""Mamba-2 state-space duality (SSD) scan — chunked, differentiable (reference ``ssm/ops/mamba2/ssd_*.py``, 2.7 kLoC Triton).

    h_t = exp(A·dt_t) h_{t-1} + dt_t · B_t ⊗ x_t ,      y_t = C_t · h_t + D x_t

Chunked evaluation (chunk length ``L``): inside a chunk the recurrence is a masked (semiseparable) matmul —
GEMM-shaped work for the tensor cores; across chunks only the ``[heads, headdim, d_state]`` states are passed.
``ssd_chunk_scan`` is the batched-GEMM formulation (einsum → cuBLAS/tcgen05 batched GEMMs); ``ssd_step`` is the
single-token recurrence used for decoding; ``ssd_reference`` is the O(l) sequential definition used by the tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def _segsum(a: torch.Tensor) -> torch.Tensor:
    """a [..., L] → S[..., i, j] = sum_{j<k<=i} a_k  (−inf above the diagonal)."""
    L = a.shape[-1]
    cs = torch.cumsum(a, dim=-1)
    s = cs[..., :, None] - cs[..., None, :]
    mask = torch.tril(torch.ones(L, L, dtype=torch.bool, device=a.device), diagonal=0)
    return s.masked_fill(~mask, float("-inf"))


def ssd_chunk_scan(x, dt, A, B, C, chunk_size: int = 128, D: Optional[torch.Tensor] = None, initial_states: Optional[torch.Tensor] = None,
                   return_final_states: bool = False):
    """x [b, l, h, p]; dt [b, l, h] (already softplus'ed, > 0); A [h] (< 0); B, C [b, l, g, n] with h % g == 0.
    Returns y [b, l, h, p] (and the final state [b, h, p, n])."""
    b, l, h, p = x.shape
    g, n = B.shape[2], B.shape[3]
    in_dtype = x.dtype
    pad = (-l) % chunk_size
    if pad:
        x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, pad))
        dt = torch.nn.functional.pad(dt, (0, 0, 0, pad))
        B = torch.nn.functional.pad(B, (0, 0, 0, 0, 0, pad))
        C = torch.nn.functional.pad(C, (0, 0, 0, 0, 0, pad))
    lp = l + pad
    c = lp // chunk_size
    rep = h // g
    xf = x.float() * dt.float().unsqueeze(-1)                       # dt-weighted input
    a = (A.float().view(1, 1, h) * dt.float())                      # [b, lp, h]  log-decay per step
    xf = xf.view(b, c, chunk_size, h, p)
    a = a.view(b, c, chunk_size, h).permute(0, 3, 1, 2)             # [b, h, c, L]
    Bf = B.float().view(b, c, chunk_size, g, n).repeat_interleave(rep, dim=3)  # [b, c, L, h, n]
    Cf = C.float().view(b, c, chunk_size, g, n).repeat_interleave(rep, dim=3)
    a_cs = torch.cumsum(a, dim=-1)                                  # [b, h, c, L]
    # 1. intra-chunk (diagonal blocks)
    Lm = torch.exp(_segsum(a))                                      # [b, h, c, L, L]
    y_diag = torch.einsum("bclhn,bcshn,bhcls,bcshp->bclhp", Cf, Bf, Lm, xf)
    # 2. state produced by each chunk
    decay_states = torch.exp(a_cs[..., -1:] - a_cs)                 # [b, h, c, L]
    states = torch.einsum("bclhn,bhcl,bclhp->bchpn", Bf, decay_states, xf)
    # 3. inter-chunk recurrence on the chunk boundary states
    if initial_states is None:
        initial_states = torch.zeros(b, h, p, n, dtype=torch.float32, device=x.device)
    if x.is_cuda:
        # O(c) scan in one kernel (csrc/extra_kernels.cu: ssd_state_fwd / reverse scan in backward) instead of the O(c²) segment-sum matmul
        from ...ops import ssd_state_passing

        prev_states, final_state = ssd_state_passing(states, a_cs[..., -1], initial_states)
    else:
        states = torch.cat([initial_states.float().unsqueeze(1), states], dim=1)   # [b, c+1, h, p, n]
        chunk_decay = torch.nn.functional.pad(a_cs[..., -1], (1, 0))               # [b, h, c+1]
        decay_chunk = torch.exp(_segsum(chunk_decay))                              # [b, h, c+1, c+1]
        new_states = torch.einsum("bhzc,bchpn->bzhpn", decay_chunk, states)
        prev_states, final_state = new_states[:, :-1], new_states[:, -1]
    # 4. contribution of the carried-in state to each position
    y_off = torch.einsum("bclhn,bchpn,bhcl->bclhp", Cf, prev_states, torch.exp(a_cs))
    y = (y_diag + y_off).reshape(b, lp, h, p)[:, :l]
    if D is not None:
        y = y + x[:, :l].float() * D.float().view(1, 1, h, -1)
    y = y.to(in_dtype)
    return (y, final_state) if return_final_states else y


def ssd_step(x, dt, A, B, C, state, D: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One decode step.  x [b, h, p]; dt [b, h]; B, C [b, g, n]; state [b, h, p, n] → (y [b, h, p], new state)."""
    h, g = x.shape[1], B.shape[1]
    if x.is_cuda and not torch.is_grad_enabled() and (D is None or D.numel() == h):
        from ...ops import ssd_step as _step_kernel       # fused in-place update + readout (csrc/extra_kernels.cu: ssd_step_kernel)

        state = state if (state.dtype == torch.float32 and state.is_contiguous()) else state.float().contiguous()
        return _step_kernel(state, x, dt, A, B, C, D), state
    rep = h // g
    Bf, Cf = B.float().repeat_interleave(rep, dim=1), C.float().repeat_interleave(rep, dim=1)
    dA = torch.exp(A.float().view(1, h) * dt.float())               # [b, h]
    dBx = torch.einsum("bh,bhn,bhp->bhpn", dt.float(), Bf, x.float())
    state = state.float() * dA[..., None, None] + dBx
    y = torch.einsum("bhpn,bhn->bhp", state, Cf)
    if D is not None:
        y = y + x.float() * D.float().view(1, h, -1)
    return y.to(x.dtype), state


def ssd_reference(x, dt, A, B, C, D=None, initial_states=None):
    """Sequential definition (tests)."""
    b, l, h, p = x.shape
    n = B.shape[-1]
    state = torch.zeros(b, h, p, n, dtype=torch.float32, device=x.device) if initial_states is None else initial_states.float()
    ys = []
    for t in range(l):
        y, state = ssd_step(x[:, t], dt[:, t], A, B[:, t], C[:, t], state, D)
        ys.append(y)
    return torch.stack(ys, dim=1), state


def causal_conv1d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, activation: Optional[str] = "silu",
                  initial_state: Optional[torch.Tensor] = None, return_final_state: bool = False):
    """Depthwise causal conv.  x [b, d, l]; weight [d, k]; state [b, d, k-1] (reference ``ssm/ops/common/causal_conv1d_triton.py``)."""
    d, k = weight.shape
    if x.is_cuda and 2 <= k <= 4 and weight.dtype == x.dtype and activation in (None, "silu", "swish"):
        from ...ops import causal_conv1d as _conv_kernel   # csrc/extra_kernels.cu: conv1d_fwd / conv1d_bwd

        xc = x.contiguous()
        y = _conv_kernel(xc, weight, bias, initial_state.to(x.dtype) if initial_state is not None else None, silu=activation is not None)
        if return_final_state:
            lf = initial_state.to(x.dtype) if initial_state is not None else x.new_zeros(x.shape[0], d, k - 1)
            return y, torch.cat([lf, xc], dim=-1)[..., -(k - 1):].contiguous()
        return y
    left = initial_state if initial_state is not None else x.new_zeros(x.shape[0], d, k - 1)
    xp = torch.cat([left.to(x.dtype), x], dim=-1)
    y = torch.nn.functional.conv1d(xp, weight.unsqueeze(1), bias, groups=d)
    if activation in ("silu", "swish"):
        y = torch.nn.functional.silu(y)
    if return_final_state:
        return y, xp[..., -(k - 1):].contiguous()
    return y


def causal_conv1d_update(x: torch.Tensor, conv_state: torch.Tensor, weight: torch.Tensor, bias=None, activation: Optional[str] = "silu"):
    """Decode step: x [b, d]; conv_state [b, d, k-1] is shifted in place."""
    window = torch.cat([conv_state, x.unsqueeze(-1)], dim=-1)       # [b, d, k]
    y = (window * weight.unsqueeze(0)).sum(-1)
    if bias is not None:
        y = y + bias
    conv_state.copy_(window[..., 1:])
    return torch.nn.functional.silu(y) if activation in ("silu", "swish") else y
