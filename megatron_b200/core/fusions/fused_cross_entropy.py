"""Vocab-parallel cross entropy, fused (reference ``fusions/fused_cross_entropy.py:12-148``) → ``ops.vocab_parallel_cross_entropy``
(``csrc/cross_entropy.cu``: one pass for max / sum-exp / target logit, ONE all-reduce, in-place backward)."""
from ... import ops


def fused_vocab_parallel_cross_entropy(vocab_parallel_logits, target, tp_group=None):
    from ..tensor_parallel.cross_entropy import vocab_parallel_cross_entropy

    return vocab_parallel_cross_entropy(vocab_parallel_logits, target, tp_group=tp_group)
