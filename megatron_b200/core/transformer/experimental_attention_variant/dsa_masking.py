"""Mask construction for DeepSeek sparse attention (reference ``experimental_attention_variant/dsa_masking.py:36-559``).

Two descriptions of "which keys may query row q see" are used side by side:

* an **additive mask** (0 / -inf, [sq, sk] or [b, sq, sk]) — arbitrary patterns;
* **row bounds** ``starts[q] <= key_position < ends[q]`` — causal + packed sequences (THD) in O(sq) memory; this is the form the
  fused kernels take (the native band-mask attention of this framework uses the same [start, end) convention).

Key positions are explicit so that the helpers also work on keys that were gathered across context-parallel ranks in a
different order than their global positions."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

_INT = (torch.int32, torch.int64)


def build_causal_mask_from_positions(query_pos: torch.Tensor, key_pos: torch.Tensor) -> torch.Tensor:
    """Additive fp32 mask [sq, sk]: -inf where ``key_pos > query_pos``."""
    assert query_pos.dtype in _INT and key_pos.dtype in _INT and query_pos.device == key_pos.device
    m = torch.zeros(query_pos.numel(), key_pos.numel(), dtype=torch.float32, device=query_pos.device)
    return m.masked_fill_(key_pos.view(1, -1) > query_pos.view(-1, 1), float("-inf"))


def generate_varlen_mask_params(cu_seqlens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row bounds of a packed causal batch: row q (global token index) sees [start of its sequence, q]."""
    assert cu_seqlens.dim() == 1 and cu_seqlens.numel() >= 2
    cu = cu_seqlens.to(torch.int64)
    q = torch.arange(int(cu[-1]), dtype=torch.int64, device=cu.device)
    return cu[torch.bucketize(q, cu, right=True) - 1], q + 1


def generate_varlen_mask_params_for_positions(cu_seqlens: torch.Tensor, query_positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same, for an arbitrary subset / order of query positions (a context-parallel rank's local queries)."""
    assert cu_seqlens.dim() == 1 and cu_seqlens.numel() >= 2 and query_positions.dtype in _INT
    cu = cu_seqlens.to(device=query_positions.device, dtype=torch.int64)
    q = query_positions.to(torch.int64)
    return cu[torch.bucketize(q, cu[1:], right=True)], q + 1


def build_valid_mask_from_starts_ends(starts: torch.Tensor, ends: torch.Tensor, key_positions: torch.Tensor) -> torch.Tensor:
    """bool [sq, sk], True = visible."""
    assert starts.dim() == 1 and starts.shape == ends.shape and key_positions.dim() == 1
    assert starts.dtype in _INT and ends.dtype in _INT and key_positions.dtype in _INT
    k = key_positions.to(torch.int64).view(1, -1)
    return (k >= starts.to(torch.int64).view(-1, 1)) & (k < ends.to(torch.int64).view(-1, 1))


def apply_starts_ends_mask_to_scores(scores: torch.Tensor, starts, ends, key_positions) -> torch.Tensor:
    """-inf outside the bounds; scores [b, sq, sk] or [b, heads, sq, sk]."""
    if scores.dim() not in (3, 4):
        raise ValueError(f"Unsupported scores ndim={scores.dim()}, expected 3 or 4.")
    valid = build_valid_mask_from_starts_ends(starts, ends, key_positions)
    return scores.masked_fill(~valid.view((1,) * (scores.dim() - 2) + valid.shape), float("-inf"))


def sort_topk_by_index(topk_indices: torch.Tensor, valid_mask: torch.Tensor, *, sk: int, topk_scores: Optional[torch.Tensor] = None,
                       invalid_score: float = float("-inf")):
    """Order the selected keys by position (sparse kernels walk the KV cache monotonically); invalid slots go last as -1."""
    if valid_mask.dtype != torch.bool or valid_mask.shape != topk_indices.shape:
        raise ValueError("valid_mask must be boolean and match topk_indices")
    if topk_scores is not None and topk_scores.shape != topk_indices.shape:
        raise ValueError("topk_scores must match topk_indices")
    order = torch.where(valid_mask, topk_indices, topk_indices.new_full((), sk)).argsort(-1)
    ok = valid_mask.gather(-1, order)
    idx = topk_indices.gather(-1, order).masked_fill(~ok, -1).contiguous()
    if topk_scores is None:
        return idx, None
    return idx, topk_scores.gather(-1, order).masked_fill(~ok, invalid_score).contiguous()


def _check(logits, valid_mask, who):
    if not logits.is_floating_point():
        raise TypeError(f"{who} expects a floating-point tensor")
    if logits.shape != valid_mask.shape:
        raise ValueError("logits and valid_mask must have the same shape")


def _shifted(logits, valid_mask, dim):
    x = logits.masked_fill(~valid_mask, torch.finfo(logits.dtype).min)
    mx = x.amax(dim, keepdim=True)
    return x - torch.where(valid_mask.any(dim, keepdim=True), mx, torch.zeros_like(mx))


def masked_softmax(logits: torch.Tensor, valid_mask: torch.Tensor, *, dim: int = -1, eps: float = 1e-10) -> torch.Tensor:
    """Softmax over the valid entries; invalid entries and fully-masked rows give exact zeros (never NaN)."""
    _check(logits, valid_mask, "masked_softmax")
    e = _shifted(logits, valid_mask, dim).exp().masked_fill(~valid_mask, 0.0)
    return (e / e.sum(dim, keepdim=True).clamp_min(eps)).masked_fill(~valid_mask, 0.0)


def masked_softmax_inplace(logits: torch.Tensor, valid_mask: torch.Tensor, *, dim: int = -1, eps: float = 1e-10) -> torch.Tensor:
    """Same result written into ``logits`` (the [b, sq, sk] score tensor is the largest activation of the unfused path)."""
    _check(logits, valid_mask, "masked_softmax_inplace")
    inv = ~valid_mask
    logits.masked_fill_(inv, torch.finfo(logits.dtype).min)
    mx = logits.amax(dim, keepdim=True)
    logits.sub_(torch.where(valid_mask.any(dim, keepdim=True), mx, torch.zeros_like(mx))).exp_().masked_fill_(inv, 0.0)
    return logits.div_(logits.sum(dim, keepdim=True).clamp_min(eps)).masked_fill_(inv, 0.0)


def masked_log_softmax(logits: torch.Tensor, valid_mask: torch.Tensor, *, dim: int = -1) -> torch.Tensor:
    """Log-softmax over the valid entries, 0 at invalid ones (so that ``target * log_prob`` sums need no second mask)."""
    _check(logits, valid_mask, "masked_log_softmax")
    x = _shifted(logits, valid_mask, dim)
    lse = x.exp().masked_fill(~valid_mask, 0.0).sum(dim, keepdim=True).clamp_min(torch.finfo(logits.dtype).tiny).log()
    return (x - lse).masked_fill(~valid_mask, 0.0)


def prepare_additive_mask(mask: Optional[torch.Tensor], *, sq: int, sk: int, b: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(score mask [sq, sk] | [b, sq, sk], attention view [1|b, 1, sq, sk], indexer view [1|b, sq, sk], valid [b, sq, sk] bool);
    ``None`` = standard causal."""
    if mask is None:
        mask = torch.full((sq, sk), float("-inf"), dtype=torch.float32, device=device).triu_(1)
    if mask.dim() == 2:
        assert mask.shape == (sq, sk), f"mask {tuple(mask.shape)} != ({sq}, {sk})"
        attn, idx = mask.view(1, 1, sq, sk), mask.view(1, sq, sk)
    elif mask.dim() == 3:
        assert mask.shape == (b, sq, sk), f"mask {tuple(mask.shape)} != ({b}, {sq}, {sk})"
        attn, idx = mask.view(b, 1, sq, sk), mask
    else:
        raise ValueError(f"mask must be [sq, sk] or [b, sq, sk], got {tuple(mask.shape)}")
    return mask, attn, idx, torch.isfinite(idx).expand(b, sq, sk)


def normalize_varlen_bounds(*, mask, varlen_starts, varlen_ends, key_positions, sk: int, device):
    """A call site passes EITHER an additive mask OR row bounds.  Bounds come back as int64 tensors on ``device``;
    ``key_positions`` stays ``None`` when the keys are in natural order (fused kernels test for that)."""
    if (varlen_starts is None) != (varlen_ends is None):
        raise ValueError("varlen_starts and varlen_ends must be given together")
    if varlen_starts is None:
        return None, None, None
    if mask is not None:
        raise ValueError("pass either an additive mask or varlen bounds, not both")
    kp = None if key_positions is None else key_positions.to(device=device, dtype=torch.int64)
    if kp is not None and kp.numel() != sk:
        raise ValueError(f"key_positions has {kp.numel()} entries for {sk} keys")
    return varlen_starts.to(device=device, dtype=torch.int64), varlen_ends.to(device=device, dtype=torch.int64), kp


def apply_sparse_validity_to_index_mask(index_mask: torch.Tensor, *, row_mask: Optional[torch.Tensor], varlen_starts, varlen_ends, key_positions) -> torch.Tensor:
    """Combine the top-k selection mask (0 at selected keys, -inf elsewhere; [b, sq, sk]) with the causal / packed visibility."""
    if varlen_starts is not None:
        sk = index_mask.shape[-1]
        s, e, kp = normalize_varlen_bounds(mask=None, varlen_starts=varlen_starts, varlen_ends=varlen_ends, key_positions=key_positions, sk=sk, device=index_mask.device)
        kp = kp if kp is not None else torch.arange(sk, device=index_mask.device)
        return apply_starts_ends_mask_to_scores(index_mask, s, e, kp)
    return index_mask if row_mask is None else index_mask + row_mask


def normalize_query_valid_rows(query_valid_rows: Optional[torch.Tensor], *, b: int, sq: int, device) -> Optional[torch.Tensor]:
    """bool [b, sq]: which query rows are real tokens (not padding between packed sequences)."""
    if query_valid_rows is None:
        return None
    v = query_valid_rows.to(device=device, dtype=torch.bool)
    if v.dim() == 1:
        v = v.view(1, -1).expand(b, -1)
    if v.shape != (b, sq):
        raise ValueError(f"query_valid_rows {tuple(v.shape)} != ({b}, {sq})")
    return v


def extract_query_valid_rows_from_packed_seq_params(packed_seq_params, sq: int, device) -> Optional[torch.Tensor]:
    """With padded packing (``cu_seqlens_q_padded``) the rows between a sequence's real end and its padded end are not tokens."""
    if packed_seq_params is None:
        return None
    cu, cup = getattr(packed_seq_params, "cu_seqlens_q", None), getattr(packed_seq_params, "cu_seqlens_q_padded", None)
    if cu is None or cup is None:
        return None
    cu, cup = cu.to(device=device, dtype=torch.int64), cup.to(device=device, dtype=torch.int64)
    q = torch.arange(sq, device=device)
    seq = (torch.bucketize(q, cup, right=True) - 1).clamp_(0, cu.numel() - 2)
    return (q - cup[seq]) < (cu[seq + 1] - cu[seq])
