"""QK-clip (MuonClip) — reference ``optimizer/qk_clip.py``: after the optimizer step, heads whose maximum attention logit exceeded the threshold
``τ`` get their query and key projection rows scaled by ``γ^α`` / ``γ^(1-α)`` with ``γ = τ / max_logit`` so the logit is pulled back to ``τ``
without touching the other heads.  The attention module records the per-head maximum logit of the step (``max_attention_logit``, reduced with MAX
over DP / TP-duplicated heads by the caller); this module applies the rescale to the fused QKV weight ``[g·(r+2)·d, h]``."""
from __future__ import annotations

from typing import Optional

import torch


def clip_qk_(linear_qkv_weight: torch.Tensor, max_logits: torch.Tensor, threshold: float, num_query_groups: int, heads_per_group: int, head_dim: int,
             alpha: float = 0.5, qkv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place.  ``max_logits [g·r]`` (local heads).  Returns the per-head factor γ (1 where nothing was clipped)."""
    g, r, d = num_query_groups, heads_per_group, head_dim
    gamma = (threshold / max_logits.float().clamp(min=1e-12)).clamp(max=1.0)                      # [g*r]
    w = linear_qkv_weight.view(g, (r + 2) * d, -1)
    gq = gamma.view(g, r)
    q_scale = gq.pow(alpha).repeat_interleave(d, dim=1)                                            # [g, r*d]
    # the key head is shared by the r query heads of its group: it can only absorb the smallest factor of the group; the rest goes to the queries
    k_gamma = gq.min(dim=1, keepdim=True).values
    k_scale = k_gamma.pow(1.0 - alpha)
    q_fix = (gq / (gq.pow(alpha) * k_scale)).repeat_interleave(d, dim=1)                          # makes q_scale·k_scale == γ for every head
    with torch.no_grad():
        w[:, : r * d].mul_((q_scale * q_fix).unsqueeze(-1).to(w.dtype))
        w[:, r * d : (r + 1) * d].mul_(k_scale.unsqueeze(-1).to(w.dtype))
        if qkv_bias is not None:
            b = qkv_bias.view(g, (r + 2) * d)
            b[:, : r * d].mul_((q_scale * q_fix).to(b.dtype))
            b[:, r * d : (r + 1) * d].mul_(k_scale.to(b.dtype))
    return gamma


def apply_qk_clip(model, threshold: float, alpha: float = 0.5) -> int:
    """Walk the attention modules that recorded ``max_attention_logit`` this step and clip them; returns the number of clipped heads."""
    n = 0
    for m in model.modules():
        logits = getattr(m, "max_attention_logit", None)
        qkv = getattr(m, "linear_qkv", None)
        if logits is None or qkv is None:
            continue
        gamma = clip_qk_(qkv.weight.data, logits, threshold, m.num_query_groups_per_partition, m.num_attention_heads_per_partition // m.num_query_groups_per_partition,
                         m.hidden_size_per_attention_head, alpha, qkv.bias.data if getattr(qkv, "bias", None) is not None else None)
        n += int((gamma < 1.0).sum())
        m.max_attention_logit = None
    return n
