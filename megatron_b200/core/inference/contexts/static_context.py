"""Static-batch KV-cache bookkeeping (reference ``inference/contexts/static_context.py``): one ``[max_seq, max_batch, ...]`` cache per layer in
``key_value_memory_dict``, a sequence offset that advances with every forward, prefill / decode mode."""
from __future__ import annotations

from ...inference_params import InferenceParams
from .base_context import BaseInferenceContext


class StaticInferenceContext(InferenceParams, BaseInferenceContext):
    def __init__(self, max_batch_size: int, max_sequence_length: int, materialize_only_last_token_logits: bool = False):
        InferenceParams.__init__(self, max_batch_size, max_sequence_length)
        BaseInferenceContext.__init__(self, materialize_only_last_token_logits)

    @classmethod
    def from_config(cls, config) -> "StaticInferenceContext":
        return cls(config.inference_max_requests, config.inference_max_seq_length)

    def is_static_batching(self) -> bool:
        return True

    def is_decode_only(self) -> bool:
        return self.decode_mode

    def __str__(self) -> str:
        return (f"StaticInferenceContext(max_seq_len = {self.max_sequence_length}, max_batch_size = {self.max_batch_size}, "
                f"sequence_len_offset = {self.sequence_len_offset}, batch_size_offset = {self.batch_size_offset}, keys = {list(self.key_value_memory_dict)})")

    def __eq__(self, other) -> bool:
        if not isinstance(other, StaticInferenceContext):
            return False
        if (self.max_sequence_length, self.max_batch_size, self.sequence_len_offset, self.batch_size_offset) != (
                other.max_sequence_length, other.max_batch_size, other.sequence_len_offset, other.batch_size_offset):
            return False
        if self.key_value_memory_dict.keys() != other.key_value_memory_dict.keys():
            return False
        import torch

        return all(all(torch.equal(a, b) for a, b in zip(self.key_value_memory_dict[k], other.key_value_memory_dict[k])) for k in self.key_value_memory_dict)

    __hash__ = None
