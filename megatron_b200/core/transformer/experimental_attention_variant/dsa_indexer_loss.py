"""KL objective that trains the DSA lightning indexer towards the main attention distribution
(reference ``experimental_attention_variant/dsa_indexer_loss.py:1-83``).

``target`` — attention probabilities of the main branch summed over heads (non-negative, any scale), ``predict_log_probs`` — the
indexer's log-softmax over the same keys.  The loss is ``KL(normalise(target) || predict)`` per query row, averaged over the
VALID rows (padding rows of packed batches carry no loss) unless the trainer normalises per token itself."""
from typing import Optional

import torch

INDEXER_LOSS_EPS = 1e-10


def normalize_indexer_target(target: torch.Tensor) -> torch.Tensor:
    return target / target.sum(-1, keepdim=True).clamp_min(INDEXER_LOSS_EPS)


def normalize_indexer_target_(target: torch.Tensor) -> torch.Tensor:
    return target.div_(target.sum(-1, keepdim=True).clamp_min(INDEXER_LOSS_EPS))


def _kl_elements(target, predict_log_probs, valid_mask):
    # 0 * log 0 = 0 through the clamp; masked positions contribute nothing even when predict_log_probs is -inf there
    t = target * (target.clamp_min(INDEXER_LOSS_EPS).log() - predict_log_probs)
    return t if valid_mask is None else torch.where(valid_mask, t, torch.zeros_like(t))


def indexer_kl_per_row(target, predict_log_probs, valid_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _kl_elements(target, predict_log_probs, valid_mask).sum(-1)


def indexer_kl_sum(target, predict_log_probs, valid_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _kl_elements(target, predict_log_probs, valid_mask).sum()


def reduce_indexer_kl_sum(kl_sum: torch.Tensor, *, num_rows: int, calculate_per_token_loss: bool, valid_row_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-token-loss runs keep the SUM (the trainer divides by the global token count); otherwise the mean over valid rows."""
    if calculate_per_token_loss:
        return kl_sum
    if valid_row_count is not None:
        return kl_sum / valid_row_count.to(device=kl_sum.device, dtype=torch.float32).clamp_min(1.0)
    return kl_sum / max(int(num_rows), 1)


def indexer_loss_from_target(target, predict_log_probs, loss_coeff: float, query_valid_rows: Optional[torch.Tensor] = None,
                             calculate_per_token_loss: bool = False, valid_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = indexer_kl_per_row(target, predict_log_probs, valid_mask)
    count = None
    if query_valid_rows is not None:
        w = query_valid_rows.to(device=rows.device, dtype=torch.float32)
        rows, count = rows * w, w.sum()
    return loss_coeff * reduce_indexer_kl_sum(rows.sum(), num_rows=rows.numel(), calculate_per_token_loss=calculate_per_token_loss, valid_row_count=count)
