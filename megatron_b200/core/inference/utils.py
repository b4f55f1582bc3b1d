"""Small helpers the serving code shares (reference ``inference/utils.py``)."""
import enum
import itertools

import torch


class InferenceMode(enum.Enum):
    TRAINING = "training"
    INFERENCE = "inference"


class Counter:
    """Monotonic id source: ``next(counter)``; ``reset()`` starts over."""

    def __init__(self, start: int = 0):
        self._start = start
        self._it = itertools.count(start)

    def __next__(self) -> int:
        return next(self._it)

    def reset(self) -> None:
        self._it = itertools.count(self._start)


def get_attention_mask(seq_length: int) -> torch.Tensor:
    """[1, 1, s, s] boolean mask, True above the diagonal (= masked)."""
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return torch.triu(torch.ones(1, 1, seq_length, seq_length, dtype=torch.bool, device=dev), diagonal=1)


def device_memory_summary() -> str:
    if not torch.cuda.is_available():
        return "no CUDA device"
    free, total = torch.cuda.mem_get_info()
    return f"allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB, free {free / 2**30:.2f} of {total / 2**30:.2f} GiB"
