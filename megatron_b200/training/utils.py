"""Training-loop utilities (reference ``megatron/training/utils.py``): masks / position ids, loss averaging, parameter norms, memory report."""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..core import parallel_state as ps
from ..core.utils import unwrap_model


def get_ltor_masks_and_position_ids(data: torch.Tensor, eod_token: Optional[int], reset_position_ids: bool = False, reset_attention_mask: bool = False,
                                    eod_mask_loss: bool = False, pad_token: Optional[int] = None, create_attention_mask: bool = True
                                    ) -> Tuple[Optional[torch.Tensor], torch.Tensor, torch.Tensor]:
    """Left-to-right masks for a token batch ``[b, s]`` → (attention_mask [b|1, 1, s, s] bool, True = masked; loss_mask [b, s]; position_ids [b, s]).

    Document resets are vectorised: a document id per token is the running count of EOD tokens, positions restart with
    ``position - position_of_document_start`` and cross-document attention is masked by comparing document ids (no Python loop
    over EOD positions, so this stays cheap for 8k-token packed samples)."""
    b, s = data.shape
    dev = data.device
    loss_mask = torch.ones(b, s, dtype=torch.float, device=dev)
    if eod_mask_loss and eod_token is not None:
        loss_mask[data == eod_token] = 0.0
    if pad_token is not None:
        loss_mask[data == pad_token] = 0.0
    pos = torch.arange(s, device=dev).unsqueeze(0).expand(b, s)
    doc = None
    if (reset_position_ids or reset_attention_mask) and eod_token is not None:
        is_eod = (data == eod_token).long()
        doc = torch.cumsum(is_eod, dim=1) - is_eod          # the EOD token still belongs to the document it ends
    if reset_position_ids and doc is not None:
        start = torch.zeros_like(pos)
        nxt = torch.where(torch.roll(data == eod_token, 1, dims=1), pos, torch.zeros_like(pos))
        nxt[:, 0] = 0
        start = torch.cummax(nxt, dim=1).values
        pos = pos - start
    att = None
    if create_attention_mask:
        causal = torch.ones(s, s, dtype=torch.bool, device=dev).triu(1)
        if reset_attention_mask and doc is not None:
            att = (causal[None] | (doc[:, :, None] != doc[:, None, :]))[:, None]
        else:
            att = causal[None, None]
    return att, loss_mask, pos.contiguous()


def average_losses_across_data_parallel_group(losses: Iterable[torch.Tensor]) -> torch.Tensor:
    t = torch.stack([l.detach().float().reshape(()) for l in losses])
    if dist.is_initialized() and ps.is_initialized():
        g = ps.get_data_parallel_group()
        dist.all_reduce(t, group=g)
        t /= dist.get_world_size(g)
    return t


def calc_params_l2_norm(model, force_create_fp32_copy: bool = False) -> float:
    """Global L2 norm of the parameters: TP-duplicated parameters counted once, expert parameters reduced over their own model-parallel group."""
    models = model if isinstance(model, (list, tuple)) else [model]
    dense_sq = torch.zeros((), dtype=torch.float64)
    expert_sq = torch.zeros((), dtype=torch.float64)
    tp_rank = ps.get_tensor_model_parallel_rank() if ps.is_initialized() else 0
    for m in models:
        for p in m.parameters():
            shared = not getattr(p, "tensor_model_parallel", False)
            if shared and tp_rank != 0:
                continue
            sq = p.detach().double().pow(2).sum().cpu()
            if getattr(p, "allreduce", True):
                dense_sq += sq
            else:
                expert_sq += sq
    if dist.is_initialized() and ps.is_initialized():
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        d = dense_sq.to(dev)
        dist.all_reduce(d, group=ps.get_model_parallel_group())
        e = expert_sq.to(dev)
        dist.all_reduce(e, group=ps.get_expert_tensor_model_pipeline_parallel_group())
        return math.sqrt(d.item() + e.item())
    return math.sqrt(dense_sq.item() + expert_sq.item())


def report_memory(name: str = "") -> str:
    if not torch.cuda.is_available():
        return f"[{name}] memory: no CUDA device"
    gib = 2.0 ** 30
    s = (f"[{name}] memory (GiB) | allocated: {torch.cuda.memory_allocated() / gib:.2f} | max allocated: {torch.cuda.max_memory_allocated() / gib:.2f}"
         f" | reserved: {torch.cuda.memory_reserved() / gib:.2f} | max reserved: {torch.cuda.max_memory_reserved() / gib:.2f}")
    if not dist.is_initialized() or (ps.is_initialized() and ps.get_data_parallel_rank() == 0):
        print(s, flush=True)
    return s


def print_params_min_max_norm(optimizer, iteration: int) -> List[str]:
    lines = []
    for gi, group in enumerate(optimizer.param_groups):
        for pi, p in enumerate(group["params"]):
            lines.append(f"iter {iteration} group {gi} param {pi} tp {int(getattr(p, 'tensor_model_parallel', False))} "
                         f"min {p.min().item():.6E} max {p.max().item():.6E} norm {torch.linalg.norm(p.float()).item():.6E}")
    print("\n".join(lines), flush=True)
    return lines


def is_last_rank() -> bool:
    return not dist.is_initialized() or dist.get_rank() == dist.get_world_size() - 1


def print_rank_last(msg: str) -> None:
    if is_last_rank():
        print(msg, flush=True)


def logical_and_across_model_parallel_group(flag: bool) -> bool:
    if not dist.is_initialized() or not ps.is_initialized():
        return flag
    t = torch.tensor([1 if flag else 0], device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=ps.get_model_parallel_group())
    return bool(t.item())


def reduce_max_stat_across_model_parallel_group(stat: Optional[float]) -> Optional[float]:
    if not dist.is_initialized() or not ps.is_initialized():
        return stat
    t = torch.tensor([-1.0 if stat is None else float(stat)], device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=ps.get_model_parallel_group())
    return None if t.item() < 0 else t.item()


def get_batch_on_this_cp_rank(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    from ..core.utils import get_batch_on_this_cp_rank as f

    return f(batch)


def param_is_not_shared(p) -> bool:
    return not getattr(p, "shared", False)


def unwrap(model):
    return unwrap_model(model)
