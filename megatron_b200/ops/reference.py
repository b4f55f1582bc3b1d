"""Plain-PyTorch implementations of every op in ``megatron_b200.ops``.

They serve three purposes: the CPU/Gloo execution path, the oracle the CUDA
kernels are tested against (``tests/test_ops_gpu.py``), and documentation of
the exact math of each fused kernel.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn.functional as F

# ---- GEMMs ---------------------------------------------------------------------


def gemm_nt(x, w, out_dtype=None):
    y = torch.matmul(x, w.t())
    return y if out_dtype is None else y.to(out_dtype)


def gemm_nn(gy, w):
    return torch.matmul(gy, w)


def gemm_tn(a, b, out=None, accumulate=False, out_dtype=None):
    """``aᵀ @ b`` with a:[M,N], b:[M,K] → [N,K]; optional accumulate into ``out``."""
    if out is not None:
        r = torch.matmul(a.t().to(torch.float32), b.to(torch.float32)) if out.dtype == torch.float32 else torch.matmul(a.t(), b)
        if accumulate:
            out.add_(r.to(out.dtype))
        else:
            out.copy_(r)
        return out
    r = torch.matmul(a.t(), b)
    return r if out_dtype is None else r.to(out_dtype)


# ---- norms ---------------------------------------------------------------------


def rms_norm_fwd(x, w, eps, zero_centered=False):
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    g = (w.float() + 1.0) if zero_centered else w.float()
    return (xf * rstd * g).to(x.dtype), rstd.squeeze(-1)


def rms_norm_bwd(gy, x, w, rstd, zero_centered=False):
    xf, gf = x.float(), gy.float()
    g = (w.float() + 1.0) if zero_centered else w.float()
    xhat = xf * rstd.unsqueeze(-1)
    gw = (gf * xhat).reshape(-1, x.shape[-1]).sum(0)
    gxhat = gf * g
    gx = rstd.unsqueeze(-1) * (gxhat - xhat * (gxhat * xhat).mean(-1, keepdim=True))
    return gx.to(x.dtype), gw.to(w.dtype)


def layer_norm_fwd(x, w, b, eps, zero_centered=False):
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    var = xf.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    g = (w.float() + 1.0) if zero_centered else w.float()
    y = (xf - mu) * rstd * g
    if b is not None:
        y = y + b.float()
    return y.to(x.dtype), mu.squeeze(-1), rstd.squeeze(-1)


def layer_norm_bwd(gy, x, w, mu, rstd, zero_centered=False, has_bias=True):
    xf, gf = x.float(), gy.float()
    g = (w.float() + 1.0) if zero_centered else w.float()
    xhat = (xf - mu.unsqueeze(-1)) * rstd.unsqueeze(-1)
    gw = (gf * xhat).reshape(-1, x.shape[-1]).sum(0)
    gb = gf.reshape(-1, x.shape[-1]).sum(0) if has_bias else None
    gxhat = gf * g
    gx = rstd.unsqueeze(-1) * (gxhat - gxhat.mean(-1, keepdim=True) - xhat * (gxhat * xhat).mean(-1, keepdim=True))
    return gx.to(x.dtype), gw.to(w.dtype), (gb.to(w.dtype) if gb is not None else None)


# ---- activations ---------------------------------------------------------------


def swiglu_fwd(y, bias=None, probs=None, clamp=None, offset=0.0):
    yf = y.float()
    if bias is not None:
        yf = yf + bias.float()
    a, b = yf.chunk(2, dim=-1)
    if clamp is not None:
        a = a.clamp(max=clamp)
        b = b.clamp(min=-clamp, max=clamp)
    out = F.silu(a) * (b + offset)
    if probs is not None:
        out = out * probs.float()
    return out.to(y.dtype)


def swiglu_bwd(g, y, bias=None, probs=None):
    """Returns (dy, dprobs). Closed form: d/da silu(a) = s(1 + a(1-s))."""
    yf, gf = y.float(), g.float()
    if bias is not None:
        yf = yf + bias.float()
    a, b = yf.chunk(2, dim=-1)
    s = torch.sigmoid(a)
    act = a * s
    dprobs = None
    if probs is not None:
        dprobs = (gf * act * b).sum(-1, keepdim=True).to(probs.dtype)
        gf = gf * probs.float()
    da = gf * b * s * (1 + a * (1 - s))
    db = gf * act
    return torch.cat([da, db], dim=-1).to(y.dtype), dprobs


def geglu_fwd(y, bias=None):
    yf = y.float() + (bias.float() if bias is not None else 0)
    a, b = yf.chunk(2, dim=-1)
    return (F.gelu(a, approximate="tanh") * b).to(y.dtype)


def bias_gelu_fwd(y, bias=None):
    yf = y.float() + (bias.float() if bias is not None else 0)
    return F.gelu(yf, approximate="tanh").to(y.dtype)


def squared_relu(x):
    return torch.pow(F.relu(x), 2)


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


# ---- RoPE ----------------------------------------------------------------------


def _rotate_half(x, interleaved):
    if not interleaved:
        x1, x2 = x.chunk(2, dim=-1)
        return torch.cat((-x2, x1), dim=-1)
    x1, x2 = x[..., ::2], x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).reshape(x.shape)


def rope_fwd(t, freqs, interleaved=False, mscale=1.0, conj=False):
    """t: [s, b, h, d]; freqs: [s, 1, 1, d_rot] (angles).  Rotates the first d_rot channels."""
    d_rot = freqs.shape[-1]
    tr, tp = t[..., :d_rot].float(), t[..., d_rot:]
    cos, sin = torch.cos(freqs.float()) * mscale, torch.sin(freqs.float()) * mscale
    if conj:
        sin = -sin
    out = tr * cos + _rotate_half(tr, interleaved) * sin
    return torch.cat((out.to(t.dtype), tp), dim=-1)


# ---- softmax / attention -------------------------------------------------------


def scaled_masked_softmax(x, mask, scale, causal=False):
    xf = x.float() * (scale if scale is not None else 1.0)
    if causal:
        sq, sk = x.shape[-2], x.shape[-1]
        cm = torch.ones(sq, sk, dtype=torch.bool, device=x.device).triu(diagonal=1 + sk - sq)
        xf = xf.masked_fill(cm, float("-inf"))
    if mask is not None:
        xf = xf.masked_fill(mask, -10000.0)
    return torch.softmax(xf, dim=-1).to(x.dtype)


def attention_fwd(q, k, v, causal=True, scale=None, window=None, cu_seqlens=None):
    """q: [sq, b, hq, d]; k, v: [sk, b, hk, d] (GQA when hk < hq).  Returns [sq, b, hq, dv]."""
    sq, b, hq, d = q.shape
    sk, _, hk, _ = k.shape
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    rep = hq // hk
    qf = q.permute(1, 2, 0, 3).float()
    kf = k.permute(1, 2, 0, 3).float().repeat_interleave(rep, dim=1)
    vf = v.permute(1, 2, 0, 3).float().repeat_interleave(rep, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        cm = torch.ones(sq, sk, dtype=torch.bool, device=q.device).triu(diagonal=1 + sk - sq)
        s = s.masked_fill(cm, float("-inf"))
    if window is not None and window[0] >= 0:
        idx_q = torch.arange(sq, device=q.device)[:, None] + (sk - sq)
        idx_k = torch.arange(sk, device=q.device)[None, :]
        s = s.masked_fill(idx_k < idx_q - window[0], float("-inf"))
    if cu_seqlens is not None:                            # packed sequences along the token dim: block-diagonal
        cu = cu_seqlens.to(device=q.device, dtype=torch.long)
        sid_q = torch.bucketize(torch.arange(sq, device=q.device), cu[1:], right=True).clamp_(max=cu.numel() - 2)
        sid_k = torch.bucketize(torch.arange(sk, device=q.device), cu[1:], right=True).clamp_(max=cu.numel() - 2)
        s = s.masked_fill(sid_q[:, None] != sid_k[None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf)
    return o.permute(2, 0, 1, 3).to(q.dtype).contiguous()


# ---- cross entropy -------------------------------------------------------------


def cross_entropy_fwd(logits, target, label_smoothing=0.0):
    """logits [..., V] (full vocab), target [...] → per-token loss (fp32)."""
    lf = logits.float()
    lse = torch.logsumexp(lf, dim=-1)
    picked = lf.gather(-1, target.unsqueeze(-1)).squeeze(-1)
    loss = lse - picked
    if label_smoothing > 0:
        v = lf.shape[-1]
        smooth = label_smoothing * v / (v - 1)
        mean_logprob = (lf - lse.unsqueeze(-1)).mean(-1)
        loss = (1 - smooth) * loss - smooth * mean_logprob
    return loss


# ---- optimizer -----------------------------------------------------------------


def adam_step(p32, g, m, v, lr, beta1, beta2, eps, wd, step, adamw=True, grad_scale=1.0, p_lowp=None):
    gf = g.float() * grad_scale
    if not adamw and wd != 0:
        gf = gf + wd * p32
    m.mul_(beta1).add_(gf, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gf, gf, value=1 - beta2)
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    denom = (v / bc2).sqrt_().add_(eps)
    upd = (m / bc1) / denom
    if adamw and wd != 0:
        upd = upd + wd * p32
    p32.add_(upd, alpha=-lr)
    if p_lowp is not None:
        p_lowp.copy_(p32)


def l2norm(tensors: List[torch.Tensor]) -> torch.Tensor:
    if not tensors:
        return torch.zeros((), dtype=torch.float32)
    return torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(t.float()) for t in tensors]))
