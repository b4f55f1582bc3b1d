"""Vocab-parallel cross entropy (reference ``tensor_parallel/cross_entropy.py:213``).

The math lives in ``megatron_b200.ops.vocab_parallel_cross_entropy`` — one fused
statistics pass + two small all-reduces (MAX, then SUM of [sum-exp ‖ target
logit]) instead of the reference's three, and an in-place backward.
"""
from __future__ import annotations


from ... import ops
from ..utils import get_pg_rank, get_pg_size, get_tensor_model_parallel_group_if_none


def vocab_parallel_cross_entropy(vocab_parallel_logits, target, label_smoothing: float = 0.0, tp_group=None):
    """logits ``[s, b, v/tp]``, target ``[s, b]`` → per-token loss ``[s, b]`` (fp32)."""
    group = get_tensor_model_parallel_group_if_none(tp_group)
    ws, rk = get_pg_size(group), get_pg_rank(group)
    vstart = rk * vocab_parallel_logits.shape[-1]
    return ops.vocab_parallel_cross_entropy(
        vocab_parallel_logits, target, group if ws > 1 else None, label_smoothing, vstart
    )


class VocabParallelCrossEntropy:
    """Namespace kept for API parity (``calculate_logits_max`` etc. are fused away)."""

    @staticmethod
    def apply(logits, target, label_smoothing=0.0, tp_group=None):
        return vocab_parallel_cross_entropy(logits, target, label_smoothing, tp_group)
