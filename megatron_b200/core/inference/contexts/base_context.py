"""What the model code needs from an inference context (reference ``inference/contexts/base_context.py``)."""
from __future__ import annotations

import abc


class BaseInferenceContext(abc.ABC):
    def __init__(self, materialize_only_last_token_logits: bool = True):
        self.materialize_only_last_token_logits = materialize_only_last_token_logits

    @abc.abstractmethod
    def is_static_batching(self) -> bool:
        ...

    def is_dynamic_batching(self) -> bool:
        return not self.is_static_batching()

    def increment_sequence_len_offset(self, increment: int) -> None:
        self.sequence_len_offset += increment

    def increment_batch_size_offset(self, increment: int) -> None:
        self.batch_size_offset += increment

    def reset_batch_size_offset(self) -> None:
        self.batch_size_offset = 0
