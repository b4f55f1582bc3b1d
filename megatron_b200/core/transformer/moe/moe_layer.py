"""Mixture-of-experts layer (reference ``transformer/moe/moe_layer.py:214``).

``route → dispatch → experts → combine`` with an optional shared expert.  The dispatcher is
chosen by ``config.moe_token_dispatcher_type`` (allgather | alltoall | flex = NVLink fused)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Optional, Union

import torch

from ... import parallel_state as ps
from ...utils import get_pg_rank, get_pg_size
from ..module import MegatronModule
from ..spec_utils import ModuleSpec, build_module
from ..transformer_config import TransformerConfig
from .router import TopKRouter
from .token_dispatcher import MoEAllGatherTokenDispatcher, MoEAlltoAllTokenDispatcher, MoEFlexTokenDispatcher


@dataclass
class MoESubmodules:
    experts: Union[ModuleSpec, type] = None
    shared_experts: Union[ModuleSpec, type] = None


class BaseMoELayer(MegatronModule, ABC):
    def __init__(self, config: TransformerConfig, layer_number: Optional[int] = None, pg_collection=None):
        super().__init__(config)
        self.pg_collection = pg_collection
        self.ep_group = pg_collection.ep if pg_collection is not None else ps.get_expert_model_parallel_group(check_initialized=False)
        self.expert_parallel_size = get_pg_size(self.ep_group)
        assert self.expert_parallel_size > 0
        assert config.num_moe_experts % self.expert_parallel_size == 0
        self.num_local_experts = config.num_moe_experts // self.expert_parallel_size
        off = get_pg_rank(self.ep_group) * self.num_local_experts
        self.local_expert_indices = [off + i for i in range(self.num_local_experts)]
        self.use_shared_expert = config.moe_shared_expert_intermediate_size is not None
        self.shared_expert_overlap = config.moe_shared_expert_overlap
        self.router = None
        self.experts = None
        self.shared_experts = None
        self.token_dispatcher = None
        self.layer_number = layer_number

    @abstractmethod
    def forward(self, hidden_states):
        ...

    def set_layer_number(self, layer_number: int):
        self.layer_number = layer_number
        self.router.set_layer_number(layer_number)


class MoELayer(BaseMoELayer):
    def __init__(self, config: TransformerConfig, submodules: Optional[MoESubmodules] = None, layer_number: Optional[int] = None, pg_collection=None):
        self.submodules = submodules
        super().__init__(config, layer_number, pg_collection)
        self.moe_layer_recompute = config.recompute_granularity == "selective" and "moe" in (config.recompute_modules or [])
        self.router = TopKRouter(config, pg_collection)
        kind = config.moe_token_dispatcher_type
        cls = {"allgather": MoEAllGatherTokenDispatcher, "alltoall": MoEAlltoAllTokenDispatcher, "alltoall_seq": MoEAlltoAllTokenDispatcher,
               "flex": MoEFlexTokenDispatcher}[kind]
        self.token_dispatcher = cls(self.num_local_experts, self.local_expert_indices, config=config, pg_collection=pg_collection)
        self.experts = build_module(submodules.experts, self.num_local_experts, config, pg_collection=pg_collection)
        if self.use_shared_expert:
            self.shared_experts = build_module(submodules.shared_experts, config=config, pg_collection=pg_collection)
        if layer_number is not None:
            self.router.set_layer_number(layer_number)
        self._training_dispatcher = self.token_dispatcher
        if config.inference_moe_token_dispatcher_type is not None:
            self.set_inference_dispatcher(config.inference_moe_token_dispatcher_type)

    def route(self, hidden_states):
        return self.router(hidden_states)

    def set_inference_dispatcher(self, kind: Optional[str]):
        """Swap in a static-shape serving dispatcher (``nccl`` | ``nvls``); ``None`` restores the training one."""
        if kind is None:
            self.token_dispatcher = self._training_dispatcher
            return
        from .token_dispatcher_inference import get_inference_token_dispatcher

        self.token_dispatcher = get_inference_token_dispatcher(kind)(self.num_local_experts, self.local_expert_indices, config=self.config,
                                                                     pg_collection=self.pg_collection)

    def _forward_impl(self, hidden_states):
        if self.training and self.config.tensor_model_parallel_size > 1 and not self.config.sequence_parallel:
            raise ValueError("during training, tensor parallelism for MoE requires sequence parallelism")
        if self.use_shared_expert:
            self.shared_experts.launch(hidden_states)
        probs, routing_map = self.route(hidden_states)
        d = self.token_dispatcher
        x, p = d.dispatch_preprocess(hidden_states, routing_map, probs)
        x, p = d.token_dispatch(x, p)
        x, tokens_per_expert, p = d.dispatch_postprocess(x, p)
        if self.config.moe_paged_stash and self.training:
            from .paged_stash import get_paged_stash_context

            with get_paged_stash_context(True, getattr(d, "num_valid_tokens", None)):
                out, mlp_bias = self.experts(x, tokens_per_expert, p)
        else:
            out, mlp_bias = self.experts(x, tokens_per_expert, p)
        out = d.combine_preprocess(out)
        out = d.token_combine(out)
        out = d.combine_postprocess(out)
        if self.use_shared_expert:
            out = out + self.shared_experts.join()
        return out, mlp_bias

    def forward(self, hidden_states: torch.Tensor):
        if self.moe_layer_recompute and self.training:
            from ...tensor_parallel.random import checkpoint

            return checkpoint(self._forward_impl, False, hidden_states)
        return self._forward_impl(hidden_states)
