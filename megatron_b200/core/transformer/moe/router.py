"""Token → expert routing (reference ``transformer/moe/router.py:144`` ``TopKRouter``)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.distributed as dist

from ... import parallel_state as ps
from ...tensor_parallel.mappings import reduce_from_tensor_model_parallel_region
from ...utils import get_pg_size
from ..module import MegatronModule
from ..transformer_config import TransformerConfig
from .moe_utils import (
    MoEAuxLossAutoScaler,
    apply_random_logits,
    apply_router_token_dropping,
    save_to_aux_losses_tracker,
    sinkhorn,
    switch_load_balancing_loss_func,
    topk_routing_with_score_function,
    z_loss_func,
)


class Router(ABC, MegatronModule):
    def __init__(self, config: TransformerConfig, pg_collection=None):
        super().__init__(config)
        self.num_experts = config.num_moe_experts
        self.layer_number = None
        self.pg_collection = pg_collection
        self.tp_group = getattr(pg_collection, "tp", None) if pg_collection is not None else ps.get_tensor_model_parallel_group(check_initialized=False)
        self.tp_cp_group = getattr(pg_collection, "tp_cp", None) if pg_collection is not None else ps.get_group("tp_cp", check_initialized=False)
        self.tp_dp_cp_group = getattr(pg_collection, "tp_dp_cp", None) if pg_collection is not None else ps.get_group("tp_dp_cp", check_initialized=False)
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight = torch.nn.Parameter(torch.empty((config.num_moe_experts, config.hidden_size), dtype=torch.float32, device=dev))
        if config.perform_initialization:
            config.init_method(self.weight)
        self.weight.data = self.weight.data.to(dtype=config.params_dtype)
        setattr(self.weight, "sequence_parallel", config.sequence_parallel)

    def gating(self, x: torch.Tensor) -> torch.Tensor:
        """``logits = x Wᵍᵀ`` computed in ``moe_router_dtype`` (fp32/fp64) when requested."""
        dt = {"fp32": torch.float32, "fp64": torch.float64}.get(self.config.moe_router_dtype or "", None)
        w = self.weight
        if dt is not None:
            return torch.nn.functional.linear(x.to(dt), w.to(dt))
        return torch.nn.functional.linear(x, w.to(x.dtype))

    @abstractmethod
    def routing(self, logits: torch.Tensor):
        ...

    def set_layer_number(self, layer_number: int):
        self.layer_number = layer_number


class TopKRouter(Router):
    """softmax / sigmoid scores, (group-limited) top-k, aux / seq-aux / global-aux balancing
    losses, z-loss, aux-loss-free expert bias, sinkhorn, capacity-based token dropping."""

    def __init__(self, config: TransformerConfig, pg_collection=None):
        super().__init__(config, pg_collection)
        self.topk = config.moe_router_topk
        self.routing_type = config.moe_router_load_balancing_type
        self.score_function = config.moe_router_score_function
        self.input_jitter = None
        self.router_replay = None
        if getattr(config, "moe_enable_routing_replay", False):
            from .router_replay import RouterReplay

            self.router_replay = RouterReplay()
        self.enable_expert_bias = config.moe_router_enable_expert_bias
        if self.enable_expert_bias:
            dev = self.weight.device
            self.register_buffer("local_tokens_per_expert", torch.zeros(config.num_moe_experts, dtype=torch.float32, device=dev), persistent=False)
            self.register_buffer("expert_bias", torch.zeros(config.num_moe_experts, dtype=torch.float32, device=dev))
        else:
            self.local_tokens_per_expert = None
            self.expert_bias = None

    def apply_input_jitter(self, x: torch.Tensor):
        eps = self.config.moe_input_jitter_eps
        if eps is None or not self.training:
            return x
        if self.input_jitter is None:
            self.input_jitter = torch.distributions.uniform.Uniform(torch.tensor(1.0 - eps, device=x.device), torch.tensor(1.0 + eps, device=x.device)).rsample
        return x * self.input_jitter(x.shape)

    def apply_z_loss(self, logits):
        if self.config.moe_z_loss_coeff is not None and self.training and torch.is_grad_enabled():
            coeff = self.config.moe_z_loss_coeff / get_pg_size(self.tp_cp_group)
            z = z_loss_func(logits, coeff)
            logits = MoEAuxLossAutoScaler.apply(logits, z)
            save_to_aux_losses_tracker("z_loss", z / coeff if coeff else z, self.layer_number, self.config.num_layers)
        return logits

    def _aux_loss(self, probs_for_loss, routing_map, activation, seq_info=None):
        coeff = self.config.moe_aux_loss_coeff
        coeff = coeff if not isinstance(coeff, (list, tuple)) else coeff[0]
        if not coeff or not self.training or not torch.is_grad_enabled():
            return activation
        T = routing_map.shape[0]
        tokens_per_expert = routing_map.sum(dim=0).float()
        total = T
        aggregated = probs_for_loss.sum(dim=0)
        # under SP each TP rank sees 1/tp of the tokens: reduce counts and prob mass over tp×cp
        if self.tp_cp_group is not None and get_pg_size(self.tp_cp_group) > 1:
            tokens_per_expert = tokens_per_expert.clone()
            dist.all_reduce(tokens_per_expert, group=self.tp_cp_group)
            aggregated = reduce_from_tensor_model_parallel_region(aggregated, group=self.tp_cp_group)
            total = T * get_pg_size(self.tp_cp_group)
        if self.routing_type == "global_aux_loss" and self.tp_dp_cp_group is not None and get_pg_size(self.tp_dp_cp_group) > 1:
            tokens_per_expert = tokens_per_expert.clone()
            dist.all_reduce(tokens_per_expert, group=ps.get_data_parallel_group())
            total = total * ps.get_data_parallel_world_size()
        if self.config.moe_router_fusion and total == T and probs_for_loss.dim() == 2:
            # no cross-rank probability mass to add: the fused kernel reduces the [T, E] probabilities against the counts directly (no [E] intermediate)
            loss = switch_load_balancing_loss_func(probs_for_loss.float(), tokens_per_expert, total, self.topk, self.num_experts, coeff, fused=True)
        else:
            loss = switch_load_balancing_loss_func(aggregated, tokens_per_expert, total, self.topk, self.num_experts, coeff)
        save_to_aux_losses_tracker("load_balancing_loss", loss / coeff, self.layer_number, self.config.num_layers)
        return MoEAuxLossAutoScaler.apply(activation, loss)

    def _seq_aux_loss(self, scores, routing_map, activation, seq_length: int, bsz: int):
        """Per-SEQUENCE load balancing (DeepSeek-V3; reference ``router.py:450-500``): the batch dimension is folded into the expert dimension
        (``[s·b, E] → [s, b·E]``), so the switch loss sums one term per sequence; divided by the batch size it is their mean."""
        coeff = self.config.moe_aux_loss_coeff
        coeff = coeff if not isinstance(coeff, (list, tuple)) else coeff[0]
        if not coeff or not self.training or not torch.is_grad_enabled():
            return activation
        sc = scores.reshape(seq_length, -1)
        rm = routing_map.reshape(seq_length, -1)
        tokens_per_expert = rm.sum(dim=0).float()
        total = seq_length
        aggregated = sc.sum(dim=0)
        if self.tp_cp_group is not None and get_pg_size(self.tp_cp_group) > 1:          # the sequence is split over tp x cp under SP / CP
            tokens_per_expert = tokens_per_expert.clone()
            dist.all_reduce(tokens_per_expert, group=self.tp_cp_group)
            aggregated = reduce_from_tensor_model_parallel_region(aggregated, group=self.tp_cp_group)
            total = seq_length * get_pg_size(self.tp_cp_group)
        loss = switch_load_balancing_loss_func(aggregated, tokens_per_expert, total, self.topk, self.num_experts, coeff) / bsz
        save_to_aux_losses_tracker("seq_load_balancing_loss", loss / coeff, self.layer_number, self.config.num_layers)
        return MoEAuxLossAutoScaler.apply(activation, loss)

    def routing(self, logits: torch.Tensor):
        seq_length, bsz = (logits.shape[0], logits.shape[1]) if logits.dim() == 3 else (logits.shape[0], 1)
        logits = logits.view(-1, self.num_experts)
        logits = self.apply_z_loss(logits)
        cfg = self.config
        if cfg.moe_router_force_load_balancing:
            logits = apply_random_logits(logits)
        if self.routing_type == "sinkhorn":
            assert cfg.moe_aux_loss_coeff == 0, "sinkhorn routing does not support aux loss"
            if self.training:
                with torch.no_grad():
                    norm = sinkhorn(logits.to(dtype=torch.float32))
                    _, idx = torch.topk(norm, k=self.topk, dim=1)
                act = torch.sigmoid(logits) if self.topk == 1 else torch.softmax(logits, dim=-1, dtype=torch.float32).type_as(logits)
            else:
                act = torch.sigmoid(logits) if self.topk == 1 else torch.softmax(logits, dim=-1, dtype=torch.float32).type_as(logits)
                _, idx = torch.topk(act, k=self.topk, dim=1)
            routing_map = torch.zeros_like(logits).int().scatter(1, idx, 1).bool()
            return act * routing_map, routing_map
        probs, routing_map = topk_routing_with_score_function(
            logits, self.topk, use_pre_softmax=cfg.moe_router_pre_softmax, num_groups=cfg.moe_router_num_groups, group_topk=cfg.moe_router_group_topk,
            scaling_factor=cfg.moe_router_topk_scaling_factor, score_function=self.score_function, expert_bias=self.expert_bias,
            router_replay=self.router_replay,
        )
        if cfg.moe_expert_capacity_factor is not None:
            probs, routing_map = apply_router_token_dropping(probs, routing_map, self.topk, cfg.moe_expert_capacity_factor,
                                                             cfg.moe_token_drop_policy, cfg.moe_pad_expert_input_to_capacity)
        if self.routing_type in ("aux_loss", "seq_aux_loss", "global_aux_loss") and self.training:
            if self.score_function == "softmax":
                scores = torch.softmax(logits, dim=-1, dtype=torch.float32)
            else:
                s = torch.sigmoid(logits.float())
                scores = s / (s.sum(-1, keepdim=True) + 1e-20)
            # the balancing losses look at the PLAIN top-k of the normalised scores, not at the map actually routed with (group limits, expert bias, capacity
            # drops must not leak into the balancing signal) — reference moe_utils.compute_routing_scores_for_aux_loss
            _, top_idx = torch.topk(scores, k=self.topk, dim=1)
            map_for_loss = torch.zeros_like(logits).int().scatter(1, top_idx, 1).bool()
            if self.routing_type == "seq_aux_loss":
                probs = self._seq_aux_loss(scores, map_for_loss, probs, seq_length, bsz)
            else:
                probs = self._aux_loss(scores, map_for_loss, probs)
        if self.enable_expert_bias and torch.is_grad_enabled():
            with torch.no_grad():
                self.local_tokens_per_expert += routing_map.sum(dim=0)
        return probs, routing_map

    def forward(self, input: torch.Tensor, padding_mask: Optional[torch.Tensor] = None):
        input = self.apply_input_jitter(input)
        logits = self.gating(input)
        return self.routing(logits)
