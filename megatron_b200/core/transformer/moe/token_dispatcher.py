"""MoE token dispatchers (reference ``transformer/moe/token_dispatcher.py``).

Six-phase interface shared by all dispatchers (reference :64-231):
``dispatch_preprocess → token_dispatch → dispatch_postprocess → [experts] →
combine_preprocess → token_combine → combine_postprocess``.

* ``MoEAllGatherTokenDispatcher``  all-gather tokens over TP×EP, local mask, reduce-scatter back
* ``MoEAlltoAllTokenDispatcher``   permute → all-to-all-v over EP → sort by local expert → … → inverse
* ``MoEFlexTokenDispatcher``       B200 path: one NVLink *push* kernel writes each token row straight
                                    into the destination rank's expert slab (symmetric memory, device-side
                                    count exchange, no host sync) — ``parallel.nvlink_moe``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List

import torch
import torch.distributed as dist

from ... import parallel_state as ps
from ...tensor_parallel.mappings import (
    all_to_all,
    gather_from_sequence_parallel_region,
    reduce_scatter_to_sequence_parallel_region,
)
from ...utils import get_pg_rank, get_pg_size
from ..transformer_config import TransformerConfig
from .moe_utils import permute, sort_chunks_by_idxs, unpermute


class MoETokenDispatcher(ABC):
    def __init__(self, config: TransformerConfig, pg_collection=None):
        self.config = config
        self.shared_experts = None
        if pg_collection is not None:
            self.ep_group, self.tp_group, self.tp_ep_group = pg_collection.ep, pg_collection.expt_tp, pg_collection.tp_ep
        else:
            self.ep_group = ps.get_expert_model_parallel_group(check_initialized=False)
            self.tp_group = ps.get_expert_tensor_parallel_group(check_initialized=False)
            self.tp_ep_group = ps.get_expert_tensor_and_model_parallel_group(check_initialized=False)
        self.ep_size, self.tp_size = get_pg_size(self.ep_group), get_pg_size(self.tp_group)
        self.ep_rank, self.tp_rank = get_pg_rank(self.ep_group), get_pg_rank(self.tp_group)

    @abstractmethod
    def dispatch_preprocess(self, tokens, routing_map, probs):
        ...

    @abstractmethod
    def token_dispatch(self, hidden_states, probs):
        ...

    @abstractmethod
    def dispatch_postprocess(self, hidden_states, probs):
        ...

    @abstractmethod
    def combine_preprocess(self, hidden_states):
        ...

    @abstractmethod
    def token_combine(self, hidden_states):
        ...

    @abstractmethod
    def combine_postprocess(self, hidden_states):
        ...

    def set_shared_experts(self, shared_experts):
        self.shared_experts = shared_experts


class MoEAllGatherTokenDispatcher(MoETokenDispatcher):
    def __init__(self, num_local_experts: int, local_expert_indices: List[int], config: TransformerConfig, pg_collection=None):
        super().__init__(config, pg_collection)
        self.num_local_experts = num_local_experts
        assert num_local_experts > 0
        self.local_expert_indices = local_expert_indices
        self.router_topk = config.moe_router_topk
        self.add_bias = config.add_bias_linear

    def dispatch_preprocess(self, hidden_states, routing_map, probs):
        self.hidden_shape = hidden_states.shape
        self.routing_map, self.probs_in = routing_map, probs
        return hidden_states.view(-1, self.hidden_shape[-1]), probs

    def token_dispatch(self, hidden_states, probs):
        if self.tp_size > 1 or self.ep_size > 1:
            self.routing_map = gather_from_sequence_parallel_region(self.routing_map, group=self.tp_ep_group)
            probs = gather_from_sequence_parallel_region(probs, group=self.tp_ep_group)
            hidden_states = gather_from_sequence_parallel_region(hidden_states, group=self.tp_ep_group, use_global_buffer=False)
        return hidden_states, probs

    def dispatch_postprocess(self, hidden_states, probs):
        self.hidden_shape_before_permute = hidden_states.shape
        lo, hi = self.local_expert_indices[0], self.local_expert_indices[-1]
        self.local_map = self.routing_map[:, lo : hi + 1].contiguous()
        self.local_probs = probs[:, lo : hi + 1].contiguous()
        tokens_per_expert = self.local_map.sum(dim=0).long().cpu()
        permuted, permuted_probs, self.reversed_mapping = permute(hidden_states, self.local_map, probs=self.local_probs)
        return permuted, tokens_per_expert, permuted_probs

    def combine_preprocess(self, hidden_states):
        return unpermute(hidden_states, self.reversed_mapping, restore_shape=self.hidden_shape_before_permute)

    def token_combine(self, hidden_states):
        if self.tp_size > 1 or self.ep_size > 1:
            hidden_states = reduce_scatter_to_sequence_parallel_region(hidden_states.to(self.local_probs.dtype), group=self.tp_ep_group).to(hidden_states.dtype)
        return hidden_states

    def combine_postprocess(self, hidden_states):
        return hidden_states.view(self.hidden_shape)


class MoEAlltoAllTokenDispatcher(MoETokenDispatcher):
    """Reference :375-960.  Experts are assigned contiguously: EP rank r owns global experts
    ``[r*L, (r+1)*L)``; a token copy destined to global expert e travels to rank ``e // L``."""

    def __init__(self, num_local_experts: int, local_expert_indices: List[int], config: TransformerConfig, pg_collection=None):
        super().__init__(config, pg_collection)
        self.num_local_experts = num_local_experts
        self.num_experts = config.num_moe_experts
        assert num_local_experts > 0
        self.local_expert_indices = local_expert_indices
        assert len(local_expert_indices) == num_local_experts
        for i in range(num_local_experts - 1):
            assert local_expert_indices[i] == local_expert_indices[i + 1] - 1, "local_expert_indices must be continuous"
        self.drop_and_pad = config.moe_pad_expert_input_to_capacity
        self.capacity = None
        L = num_local_experts
        # received chunk order is [src tp*ep rank][local expert]; experts want [local expert][src]
        n_src = self.ep_size * self.tp_size
        self._sort_in = torch.arange(n_src * L).reshape(-1, L).T.ravel().tolist()
        self._sort_out = torch.arange(n_src * L).reshape(L, -1).T.ravel().tolist()

    def _preprocess_counts(self, routing_map: torch.Tensor):
        """Exchange per-expert token counts; the single host sync of the dispatcher."""
        L, E = self.num_local_experts, self.num_experts
        local_counts = routing_map.sum(dim=0).long()  # [E]
        if self.drop_and_pad:
            # the SAME capacity the router's token dropping encoded in routing_map (reference token_dispatcher.py: get_capacity with the factor)
            from .moe_utils import get_capacity

            assert self.config.moe_expert_capacity_factor is not None, "moe_pad_expert_input_to_capacity requires moe_expert_capacity_factor"
            cap = get_capacity(routing_map.shape[0] * self.config.moe_router_topk, E, self.config.moe_expert_capacity_factor)
            self.capacity = cap
            self.num_out_tokens = cap * E
            self.input_splits = self.output_splits = None
            self.output_splits_tp = None
            return torch.full((L,), cap * self.tp_size * self.ep_size, dtype=torch.long)
        self.num_out_tokens = None
        if self.ep_size > 1 or self.tp_size > 1:
            gathered = torch.empty(self.ep_size * self.tp_size, E, dtype=torch.long, device=local_counts.device)
            dist.all_gather_into_tensor(gathered.view(-1), local_counts.contiguous(), group=self.tp_ep_group)
            gathered = gathered.view(self.ep_size, self.tp_size, E) if self.tp_size > 1 else gathered.view(self.ep_size, 1, E)
            # NB: tp_ep group rank order is (tp fastest) → reshape accordingly
            gathered = gathered.reshape(self.ep_size * self.tp_size, E)
            mine = gathered[:, self.local_expert_indices[0] : self.local_expert_indices[-1] + 1]  # [src, L]
            host = torch.cat([local_counts.view(-1), mine.reshape(-1)]).cpu()
            lc, mine_h = host[:E], host[E:].view(-1, L)
            self.input_splits = lc.view(self.ep_size, L).sum(dim=1).tolist()
            # all_to_all over EP only: sources with my tp rank
            if self.tp_size > 1:
                src_ep = mine_h.view(self.ep_size, self.tp_size, L)[:, self.tp_rank, :]
            else:
                src_ep = mine_h
            self.output_splits = src_ep.sum(dim=1).tolist()
            self.num_global_tokens_per_local_expert = src_ep  # [ep, L]
            tokens_per_local_expert = src_ep.sum(dim=0)
            self.output_splits_tp = None
            if self.tp_size > 1:
                all_tp = mine_h.view(self.ep_size, self.tp_size, L)
                self.output_splits_tp = all_tp.sum(dim=(0, 2)).tolist()
                self.num_global_tokens_per_local_expert = all_tp.permute(1, 0, 2).reshape(-1, L)  # [tp*ep, L]
                tokens_per_local_expert = all_tp.sum(dim=(0, 1))
            return tokens_per_local_expert
        self.input_splits = self.output_splits = None
        self.output_splits_tp = None
        self.num_global_tokens_per_local_expert = local_counts.view(1, -1).cpu()
        return local_counts.cpu()

    def dispatch_preprocess(self, hidden_states, routing_map, probs):
        self.hidden_shape = hidden_states.shape
        self.routing_map, self.probs = routing_map, probs
        assert probs.dim() == 2 and routing_map.dim() == 2 and routing_map.dtype == torch.bool
        hidden_states = hidden_states.view(-1, self.hidden_shape[-1])
        self.tokens_per_expert = self._preprocess_counts(routing_map)
        self.hidden_shape_before_permute = hidden_states.shape
        permuted, permuted_probs, self.reversed_local_input_permutation_mapping = permute(
            hidden_states, routing_map, probs=probs, num_out_tokens=self.num_out_tokens, drop_and_pad=self.drop_and_pad
        )
        return permuted, permuted_probs

    def token_dispatch(self, permuted, permuted_probs):
        if self.ep_size > 1:
            permuted = all_to_all(self.ep_group, permuted, self.output_splits, self.input_splits)
            permuted_probs = all_to_all(self.ep_group, permuted_probs, self.output_splits, self.input_splits)
        return permuted, permuted_probs

    def dispatch_postprocess(self, tokens, probs):
        if self.tp_size > 1:
            tokens = gather_from_sequence_parallel_region(tokens, group=self.tp_group, output_split_sizes=self.output_splits_tp)
            probs = gather_from_sequence_parallel_region(probs, group=self.tp_group, output_split_sizes=self.output_splits_tp)
        if self.num_local_experts > 1:
            if self.drop_and_pad:
                n_src = self.tp_size * self.ep_size
                tokens = tokens.view(n_src, self.num_local_experts, self.capacity, -1).transpose(0, 1).reshape(self.num_local_experts * n_src * self.capacity, -1)
                probs = probs.view(n_src, self.num_local_experts, self.capacity).transpose(0, 1).reshape(-1)
            else:
                tokens, probs = sort_chunks_by_idxs(tokens, self.num_global_tokens_per_local_expert.ravel(), self._sort_in, probs=probs)
        return tokens, self.tokens_per_expert, probs

    def combine_preprocess(self, hidden_states):
        if self.num_local_experts > 1:
            if self.drop_and_pad:
                n_src = self.tp_size * self.ep_size
                hidden_states = hidden_states.view(self.num_local_experts, n_src, self.capacity, -1).transpose(0, 1).reshape(n_src * self.num_local_experts * self.capacity, -1)
            else:
                hidden_states, _ = sort_chunks_by_idxs(hidden_states, self.num_global_tokens_per_local_expert.T.ravel(), self._sort_out)
        if self.tp_size > 1:
            hidden_states = reduce_scatter_to_sequence_parallel_region(hidden_states, group=self.tp_group, input_split_sizes=self.output_splits_tp)
        return hidden_states

    def token_combine(self, hidden_states):
        if self.ep_size > 1:
            hidden_states = all_to_all(self.ep_group, hidden_states, self.input_splits, self.output_splits)
        return hidden_states

    def combine_postprocess(self, permuted):
        out = unpermute(permuted, self.reversed_local_input_permutation_mapping, restore_shape=self.hidden_shape_before_permute,
                        routing_map=self.routing_map, drop_and_pad=self.drop_and_pad)
        return out.view(self.hidden_shape)


class MoEFlexTokenDispatcher(MoEAlltoAllTokenDispatcher):
    """Fused NVLink dispatch/combine (stands in for the reference's DeepEP / HybridEP / NCCL-EP
    managers, :1001-2085).  Falls back to the all-to-all path when the EP group has no
    symmetric heap (CPU, or ``collectives.enable_for_group`` not called)."""

    def __init__(self, num_local_experts, local_expert_indices, config, pg_collection=None):
        super().__init__(num_local_experts, local_expert_indices, config, pg_collection)
        self._fused = None

    def _backend(self, t):
        if not t.is_cuda or self.tp_size > 1 or self.drop_and_pad:
            return None
        from ....parallel import collectives

        be = collectives.backend_for(self.ep_group)
        return be if (be is not None and hasattr(be, "moe_dispatch")) else None

    def dispatch_preprocess(self, hidden_states, routing_map, probs):
        be = self._backend(hidden_states)
        if be is None:
            self._fused = None
            return super().dispatch_preprocess(hidden_states, routing_map, probs)
        self._fused = be
        self.hidden_shape = hidden_states.shape
        self.routing_map, self.probs = routing_map, probs
        return hidden_states.view(-1, self.hidden_shape[-1]), probs

    def token_dispatch(self, tokens, probs):
        if self._fused is None:
            return super().token_dispatch(tokens, probs)
        self._handle, out, out_probs, self.tokens_per_expert = self._fused.moe_dispatch(tokens, self.routing_map, probs, self.num_local_experts, topk=self.config.moe_router_topk)
        return out, out_probs

    def dispatch_postprocess(self, tokens, probs):
        if self._fused is None:
            return super().dispatch_postprocess(tokens, probs)
        return tokens, self.tokens_per_expert, probs  # already grouped by local expert

    def combine_preprocess(self, hidden_states):
        return hidden_states if self._fused is not None else super().combine_preprocess(hidden_states)

    def token_combine(self, hidden_states):
        if self._fused is None:
            return super().token_combine(hidden_states)
        return self._fused.moe_combine(hidden_states, self._handle)

    def combine_postprocess(self, hidden_states):
        if self._fused is None:
            return super().combine_postprocess(hidden_states)
        return hidden_states.view(self.hidden_shape)
