"""End-of-step gradient fix-ups (reference ``distributed/finalize_model_grads.py:560``):

1. wait for / launch the data-parallel bucket reductions,
2. all-reduce grads of params replicated across TP but fed by sequence-sharded
   activations (norm weights, QK-norm, row-linear biases) — ONE flattened all-reduce,
3. all-reduce tied word-embedding grads between first and last pipeline stage,
4. MoE router expert-bias update,
5. per-token loss normalisation (divide by the global token count).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors

from .. import parallel_state as ps
from ..utils import get_attr_wrapped_model, get_model_config, get_pg_size


def _grad_of(p):
    return p.main_grad if hasattr(p, "main_grad") else p.grad


def _allreduce_coalesced(grads: List[torch.Tensor], group, avg: bool = False):
    if not grads or group is None or get_pg_size(group) == 1:
        return
    flat = _flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    if avg:
        flat.div_(get_pg_size(group))
    for g, synced in zip(grads, _unflatten_dense_tensors(flat, grads)):
        g.copy_(synced)


def _allreduce_non_tensor_model_parallel_grads(model, config, tp_group):
    """SP norm weights / QK-norm / SP biases: replicated params touched by 1/tp of the tokens."""
    if get_pg_size(tp_group) <= 1:
        return
    grads = []
    for chunk in model:
        for name, p in get_attr_wrapped_model(chunk, "named_parameters")():
            if not p.requires_grad:
                continue
            sp = getattr(p, "sequence_parallel", False) and config.sequence_parallel
            qk = config.qk_layernorm and ("q_layernorm" in name or "k_layernorm" in name)
            avg = getattr(p, "average_gradients_across_tp_domain", False)
            if sp or qk or avg:
                g = _grad_of(p)
                if g is not None:
                    grads.append(g.data)
    _allreduce_coalesced(grads, tp_group)


def _allreduce_word_embedding_grads(model, config, embd_group, pp_group):
    """Tied embeddings: first and last stage each hold a copy; sum their grads."""
    if embd_group is None or get_pg_size(embd_group) <= 1 or not ps.is_rank_in_embedding_group(ignore_virtual=True):
        return
    if ps.is_pipeline_first_stage(ignore_virtual=True):
        chunk = model[0]
    elif ps.is_pipeline_last_stage(ignore_virtual=True):
        chunk = model[-1]
    else:
        chunk = model[0]
    m = get_attr_wrapped_model(chunk, "pre_process", return_model_obj=True)
    if getattr(m, "share_embeddings_and_output_weights", False):
        w = m.shared_embedding_or_output_weight()
        g = _grad_of(w)
        if g is not None:
            dist.all_reduce(g, group=embd_group)


def _allreduce_position_embedding_grads(model, config, pos_group):
    if pos_group is None or get_pg_size(pos_group) <= 1 or not ps.is_rank_in_position_embedding_group():
        return
    m = get_attr_wrapped_model(model[0], "pre_process", return_model_obj=True)
    emb = getattr(m, "embedding", None)
    pe = getattr(emb, "position_embeddings", None)
    if pe is not None:
        dist.all_reduce(_grad_of(pe.weight), group=pos_group)


def _update_router_expert_bias(model, config, group):
    """Aux-loss-free balancing: nudge per-expert bias against measured load (DeepSeek-V3)."""
    tokens, biases = [], []
    for chunk in model:
        for m in get_attr_wrapped_model(chunk, "modules")():
            if hasattr(m, "expert_bias") and getattr(m, "expert_bias", None) is not None and hasattr(m, "local_tokens_per_expert"):
                tokens.append(m.local_tokens_per_expert)
                biases.append(m.expert_bias)
    if not biases:
        return
    st = torch.stack(tokens)
    if group is not None and get_pg_size(group) > 1:
        dist.all_reduce(st, group=group)
    with torch.no_grad():
        avg = st.float().mean(dim=-1, keepdim=True)
        upd = torch.sign(avg - st.float()) * config.moe_router_bias_update_rate
        for t, b, u in zip(tokens, biases, upd):
            b.add_(u)
            t.zero_()


def finalize_model_grads(model: List[torch.nn.Module], num_tokens: Optional[torch.Tensor] = None, pg_collection=None, force_all_reduce: bool = False):
    config = get_model_config(model[0])
    if pg_collection is None:
        tp = ps.get_tensor_model_parallel_group(check_initialized=False)
        pp = ps.get_pipeline_model_parallel_group(check_initialized=False)
        embd = ps.get_embedding_group(check_initialized=False)
        pos = ps.get_position_embedding_group(check_initialized=False)
        dp_cp = ps.get_group("dp_cp", check_initialized=False)
        tp_dp_cp = ps.get_group("tp_dp_cp", check_initialized=False)
    else:
        tp, pp, embd, pos, dp_cp = pg_collection.tp, pg_collection.pp, getattr(pg_collection, "embd", None), getattr(pg_collection, "pos_embd", None), pg_collection.dp_cp
        tp_dp_cp = getattr(pg_collection, "tp_dp_cp", None)
    timers = config.timers
    if timers is not None:
        timers("all-grads-sync", log_level=1).start(barrier=config.barrier_with_L1_time)
    for chunk in model:
        chunk.finish_grad_sync(force_all_reduce=force_all_reduce)
    if timers is not None:
        timers("all-grads-sync").stop()
        timers("non-tensor-parallel-grads-all-reduce", log_level=1).start(barrier=config.barrier_with_L1_time)
    _allreduce_non_tensor_model_parallel_grads(model, config, tp)
    if timers is not None:
        timers("non-tensor-parallel-grads-all-reduce").stop()
        timers("embedding-grads-all-reduce", log_level=1).start(barrier=config.barrier_with_L1_time)
    _allreduce_word_embedding_grads(model, config, embd, pp)
    _allreduce_position_embedding_grads(model, config, pos)
    if timers is not None:
        timers("embedding-grads-all-reduce").stop()
    if config.moe_router_enable_expert_bias:
        _update_router_expert_bias(model, config, tp_dp_cp)
    if num_tokens is not None:
        # the last stage knows the token count; broadcast over pp, sum over dp×cp
        if pp is not None and get_pg_size(pp) > 1:
            dist.broadcast(num_tokens, src=ps.get_pipeline_model_parallel_last_rank(), group=pp)
        if dp_cp is not None and get_pg_size(dp_cp) > 1:
            dist.all_reduce(num_tokens, group=dp_cp)
        for chunk in model:
            if float(num_tokens) > 0:
                chunk.scale_gradients(1.0 / float(num_tokens))
