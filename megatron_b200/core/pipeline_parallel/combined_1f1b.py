"""Forward/backward-overlapped 1F1B inside one rank (reference ``pipeline_parallel/combined_1f1b.py``:
``combined_1f1b_schedule_for_no_pipelining`` :35, ``combined_forward_backward_step`` :281).

With expert parallelism the MoE all-to-all sits on the critical path of every layer.  This schedule keeps TWO micro-batches
in flight per rank — the forward of micro-batch *i* and the backward of micro-batch *i-1* — and walks them through the
layer schedule plans (``models/common/model_chunk_schedule_plan.py``) so that a communication node of one is always issued
next to a compute node of the other, on different CUDA streams.

    step 0      : F0
    step i      : F_i  ‖  B_{i-1}        (1 ≤ i < n)
    step n      :         B_{n-1}

Gradient accumulation, loss scaling and the final gradient synchronisation follow ``forward_backward_no_pipelining``.
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import Callable, Iterator, List, Optional

import torch

from ..models.common.model_chunk_schedule_plan import TransformerModelChunkSchedulePlan
from .utils import get_comm_stream, get_comp_stream, set_streams


def default_plan_builder(batch: dict, model, loss_func: Optional[Callable]):
    return TransformerModelChunkSchedulePlan(
        model, batch["tokens"], batch.get("position_ids"), batch.get("attention_mask"), labels=batch.get("labels"),
        loss_mask=batch.get("loss_mask"), packed_seq_params=batch.get("packed_seq_params"), loss_func=loss_func,
    )


def _masked_mean(loss_mask):
    def f(per_token):
        if loss_mask is None:
            return per_token.float().mean()
        m = loss_mask.reshape(-1).float()
        return (per_token.reshape(-1).float() * m).sum() / m.sum().clamp(min=1)

    return f


def combined_forward_backward_step(f_plan, b_plan, grad=None):
    """One overlapped step; either plan may be ``None``.  Returns the forward plan's loss (or ``None``)."""
    return TransformerModelChunkSchedulePlan.run(f_plan, b_plan, grad=grad)


def combined_1f1b_schedule_for_no_pipelining(*, data_iterator: Iterator, model, num_microbatches: int, plan_builder: Callable = default_plan_builder,
                                             loss_func_factory: Optional[Callable] = None, forward_only: bool = False, config=None,
                                             no_sync_func: Optional[Callable] = None, grad_scale: Optional[float] = None) -> List[torch.Tensor]:
    """Run ``num_microbatches`` micro-batches with F_i ‖ B_{i-1} overlap.  ``loss_func_factory(batch) -> f(per_token_loss) -> scalar``."""
    if get_comp_stream() is None and torch.cuda.is_available():
        set_streams()
    no_sync = no_sync_func or getattr(model, "no_sync", None) or nullcontext
    scale = (1.0 / num_microbatches) if grad_scale is None else grad_scale
    losses: List[torch.Tensor] = []

    def make(batch):
        base = loss_func_factory(batch) if loss_func_factory is not None else _masked_mean(batch.get("loss_mask"))
        return plan_builder(batch, model, lambda out: base(out) * scale)

    if forward_only:
        with torch.no_grad():
            for _ in range(num_microbatches):
                losses.append(combined_forward_backward_step(make(next(data_iterator)), None).detach() / scale)
        return losses

    prev = None
    with no_sync():
        for i in range(num_microbatches):
            cur = make(next(data_iterator))
            combined_forward_backward_step(cur, prev)
            losses.append(cur.loss.detach() / scale)
            prev = cur
    # the final backward runs outside no_sync so DDP-style gradient reduction fires on it
    combined_forward_backward_step(None, prev)
    cs = get_comm_stream()
    if cs is not None:
        torch.cuda.current_stream().wait_stream(cs)
    if config is not None and getattr(config, "finalize_model_grads_func", None) is not None:
        config.finalize_model_grads_func([model], None)
    return losses
