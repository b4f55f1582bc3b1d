"""SOAP — Adam run in the eigenbasis of Shampoo's Kronecker factors (reference ``optimizer/emerging_optimizers.py``, which wraps
the external ``emerging_optimizers`` package's SOAP; implemented here directly, after arXiv:2409.11321).

For a 2-D weight ``W [m, n]`` with gradient ``G``:

    L ← β_s L + (1-β_s) G Gᵀ        R ← β_s R + (1-β_s) Gᵀ G           (Shampoo statistics)
    Q_L, Q_R = eigenvectors of L, R, refreshed every ``f`` steps by ONE power-iteration + QR step
    G' = Q_Lᵀ G Q_R;   M ← β₁ M + (1-β₁) G;   V ← β₂ V + (1-β₂) G'²   (second moment lives in the rotated space)
    W ← W - lr · Q_L [ (Q_Lᵀ M̂ Q_R) / (√V̂ + ε) ] Q_Rᵀ - lr · wd · W

A side larger than ``soap_max_precond_dim`` keeps the identity basis (one-sided SOAP), which is what makes the embedding-sized
matrices affordable; with both sides skipped the rule IS AdamW (tested).  1-D parameters, embeddings and the output head go through
the calling optimizer's AdamW like with Muon.  The rotations are GEMMs of the weight's own size — tcgen05 work — and the basis
refresh is amortised over ``f`` steps.  Like Muon the rule needs whole matrices: pair it with ``LayerWiseDistributedOptimizer``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


def is_soap_param(param, master: Optional[torch.Tensor] = None) -> bool:
    t = master if master is not None else param
    return t.dim() == 2 and not getattr(param, "is_embedding_or_output_parameter", False) and not getattr(param, "is_router_parameter", False)


def _eigenbasis(stat: torch.Tensor) -> torch.Tensor:
    """Eigenvectors of a PSD matrix, largest eigenvalue first."""
    s = stat.double() if stat.device.type == "cpu" else stat.float()
    s = s + 1e-30 * torch.eye(s.shape[0], device=s.device, dtype=s.dtype)
    _, q = torch.linalg.eigh(s)
    return q.flip(1).to(stat.dtype)


def _refresh_basis(stat: torch.Tensor, q: torch.Tensor, v: torch.Tensor, dim: int):
    """One orthogonal-iteration step ``Q ← qr(stat · Q)``; columns are re-sorted by their Rayleigh quotients first and the second
    moment ``v`` is permuted along ``dim`` the same way so it keeps describing the same directions."""
    est = torch.einsum("ij,ik,kj->j", q, stat, q)
    order = torch.argsort(est, descending=True)
    q = q[:, order]
    v = v.index_select(dim, order)
    qn, _ = torch.linalg.qr((stat @ q).float())
    return qn.to(stat.dtype), v


def init_soap_state(master: torch.Tensor, max_precond_dim: int) -> Dict[str, torch.Tensor]:
    m, n = master.shape
    st: Dict[str, torch.Tensor] = {}
    if m <= max_precond_dim:
        st["L"] = torch.zeros(m, m, device=master.device, dtype=torch.float32)
    if n <= max_precond_dim:
        st["R"] = torch.zeros(n, n, device=master.device, dtype=torch.float32)
    return st


def _project(x, st, back=False):
    ql, qr = st.get("QL"), st.get("QR")
    if ql is not None:
        x = (ql @ x) if back else (ql.t() @ x)
    if qr is not None:
        x = (x @ qr.t()) if back else (x @ qr)
    return x


def soap_step(master: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, st: Dict[str, torch.Tensor], *, lr: float,
              beta1: float, beta2: float, eps: float, weight_decay: float, step: int, shampoo_beta: float = 0.95, precondition_frequency: int = 10,
              precondition_warmup: bool = True) -> None:
    """``step`` is 1-based.  The first call only seeds the statistics and their eigenbases (no weight update), as in the paper's code."""
    g = grad.float()
    has_basis = "QL" in st or "QR" in st or ("L" not in st and "R" not in st)
    first = not st.get("_seeded", False)
    if "L" in st:
        st["L"].mul_(0.0 if first else shampoo_beta).add_(g @ g.t(), alpha=1.0 if first else 1 - shampoo_beta)
    if "R" in st:
        st["R"].mul_(0.0 if first else shampoo_beta).add_(g.t() @ g, alpha=1.0 if first else 1 - shampoo_beta)
    if first:
        if "L" in st:
            st["QL"] = _eigenbasis(st["L"])
        if "R" in st:
            st["QR"] = _eigenbasis(st["R"])
        st["_seeded"] = True
        st["_t"] = 0
        if ("L" in st or "R" in st) and precondition_warmup:
            return
    del has_basis
    st["_t"] = t = st["_t"] + 1
    gp = _project(g, st)
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(gp, gp, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    denom = (exp_avg_sq / bc2).sqrt_().add_(eps)
    upd = _project(_project(exp_avg, st) / bc1 / denom, st, back=True)
    if weight_decay:
        master.mul_(1.0 - lr * weight_decay)
    master.add_(upd.to(master.dtype), alpha=-lr)
    if t % max(precondition_frequency, 1) == 0:
        if "L" in st:
            st["QL"], v = _refresh_basis(st["L"], st["QL"], exp_avg_sq, 0)
            exp_avg_sq.copy_(v)
        if "R" in st:
            st["QR"], v = _refresh_basis(st["R"], st["QR"], exp_avg_sq, 1)
            exp_avg_sq.copy_(v)
