"""Token sampling (reference ``inference/text_generation_controllers`` samplers): greedy, temperature,
top-k, top-p; vectorised over the batch with per-request parameters."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class SamplingParams:
    temperature: float = 1.0
    top_k: int = 0
    top_p: float = 0.0
    num_tokens_to_generate: int = 32
    return_log_probs: bool = False
    stop_token_ids: tuple = ()
    seed: Optional[int] = None


def sample(logits: torch.Tensor, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, generator: Optional[torch.Generator] = None,
           vocab_size: Optional[int] = None) -> torch.Tensor:
    """logits [b, v] → token ids [b]."""
    assert logits.dim() == 2
    if vocab_size is not None and vocab_size < logits.shape[-1]:
        logits = logits.clone()
        logits[:, vocab_size:] = float("-inf")  # padded vocab entries can never be sampled
    if top_k == 1 or temperature == 0.0:
        return torch.argmax(logits, dim=-1)
    logits = logits.float() / max(temperature, 1e-6)
    if top_k > 0:
        kth = torch.topk(logits, min(top_k, logits.shape[-1]), dim=-1).values[:, -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if 0.0 < top_p < 1.0:
        sorted_logits, idx = torch.sort(logits, dim=-1, descending=True)
        cum = torch.softmax(sorted_logits, dim=-1).cumsum(dim=-1)
        remove = cum - torch.softmax(sorted_logits, dim=-1) > top_p
        sorted_logits = sorted_logits.masked_fill(remove, float("-inf"))
        logits = torch.full_like(logits, float("-inf")).scatter(-1, idx, sorted_logits)
    probs = torch.softmax(logits, dim=-1)
    return torch.multinomial(probs, 1, generator=generator).squeeze(-1)
