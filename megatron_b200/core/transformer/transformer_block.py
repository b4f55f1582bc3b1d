"""Stack of transformer layers for one pipeline (virtual) stage
(reference ``transformer/transformer_block.py:267``)."""
from __future__ import annotations

from contextlib import nullcontext
from dataclasses import dataclass
from typing import List, Optional, Union

import torch

from .. import parallel_state as ps
from ..dist_checkpointing.mapping import ShardedStateDict
from ..tensor_parallel.random import checkpoint, get_cuda_rng_tracker
from ..utils import make_viewless_tensor
from .module import GraphableMegatronModule
from .spec_utils import ModuleSpec, build_module
from .torch_norm import FusedNorm
from .transformer_config import TransformerConfig
from .transformer_layer import BaseTransformerLayer, get_transformer_layer_offset
from .utils import sharded_state_dict_default


def get_num_layers_to_build(config: TransformerConfig, vp_stage: Optional[int] = None, pp_rank: Optional[int] = None) -> int:
    """Number of layers this (pp_rank, vp_stage) instantiates (reference :76-205)."""
    pp = config.pipeline_model_parallel_size
    vp = config.virtual_pipeline_model_parallel_size
    pp_rank = ps.get_pipeline_model_parallel_rank() if pp_rank is None else pp_rank
    first, last = config.num_layers_in_first_pipeline_stage, config.num_layers_in_last_pipeline_stage
    if pp > 1 and (first is not None or last is not None):
        mid_stages = pp - (first is not None) - (last is not None)
        n_mid = config.num_layers - (first or 0) - (last or 0)
        if pp_rank == 0 and first is not None:
            n = first
        elif pp_rank == pp - 1 and last is not None:
            n = last
        else:
            n = n_mid // mid_stages
        return n // vp if vp is not None else n
    total = config.num_layers
    if config.account_for_embedding_in_pipeline_split:
        total += 1
    if config.account_for_loss_in_pipeline_split:
        total += 1
    n = total // pp
    if vp is not None:
        n //= vp
    v = vp_stage if vp_stage is not None else (ps.get_virtual_pipeline_model_parallel_rank() or 0)
    is_first = pp_rank == 0 and (vp is None or v == 0)
    is_last = pp_rank == pp - 1 and (vp is None or v == vp - 1)
    if config.account_for_embedding_in_pipeline_split and is_first:
        n -= 1
    if config.account_for_loss_in_pipeline_split and is_last:
        n -= 1
    return n


@dataclass
class TransformerBlockSubmodules:
    layer_specs: List[ModuleSpec] = None
    layer_norm: Optional[Union[ModuleSpec, torch.nn.Module]] = None


def _get_block_submodules(config, spec, vp_stage=None, pp_rank=None) -> TransformerBlockSubmodules:
    if isinstance(spec, TransformerBlockSubmodules):
        return spec
    if isinstance(spec, ModuleSpec):
        mod = spec.module
        if isinstance(mod, type) and issubclass(mod, TransformerBlock):
            return spec.submodules
        if isinstance(mod, type) and issubclass(mod, BaseTransformerLayer):
            n = get_num_layers_to_build(config, vp_stage, pp_rank)
            return TransformerBlockSubmodules(layer_specs=[spec] * n, layer_norm=FusedNorm)
    raise Exception(f"specialize for {type(spec).__name__}")


class TransformerBlock(GraphableMegatronModule):
    def __init__(self, config: TransformerConfig, spec, post_layer_norm: bool = True, pre_process: bool = True,
                 post_process: bool = True, pg_collection=None, vp_stage: Optional[int] = None):
        super().__init__(config, vp_stage=vp_stage)
        self.submodules = _get_block_submodules(config, spec, vp_stage)
        self.post_layer_norm, self.pre_process, self.post_process = post_layer_norm, pre_process, post_process
        self.pg_collection = pg_collection
        self.input_tensor = None
        self.checkpoint_core_attention = config.recompute_granularity == "selective"
        self.num_layers_per_pipeline_rank = len(self.submodules.layer_specs)
        kw = {"pg_collection": pg_collection} if pg_collection is not None else {}
        self.layers = torch.nn.ModuleList(
            [build_module(s, config=config, layer_number=i + 1, vp_stage=vp_stage, **kw) for i, s in enumerate(self.submodules.layer_specs)]
        )
        if self.submodules.layer_norm is not None and self.post_process and self.post_layer_norm:
            self.final_layernorm = build_module(self.submodules.layer_norm, config=config, hidden_size=config.hidden_size, eps=config.layernorm_epsilon)
        else:
            self.final_layernorm = None

    def _get_layer(self, i):
        return self.layers[i]

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor

    def _checkpointed_forward(self, hidden_states, attention_mask, context, context_mask, rotary_pos_emb, attention_bias, packed_seq_params):
        """Full-layer recompute, ``uniform`` (chunks of k layers) or ``block`` (first k layers)."""

        def custom(start, end):
            def fwd(hs, am, ctx, cm, rpe):
                for i in range(start, end):
                    hs, ctx = self.layers[i](hs, attention_mask=am, context=ctx, context_mask=cm, rotary_pos_emb=rpe,
                                             attention_bias=attention_bias, inference_context=None, packed_seq_params=packed_seq_params)
                return hs, ctx

            return fwd

        k = self.config.recompute_num_layers
        n = self.num_layers_per_pipeline_rank
        dist_saved = self.config.distribute_saved_activations
        if self.config.recompute_method == "uniform":
            i = 0
            while i < n:
                hidden_states, context = checkpoint(custom(i, min(i + k, n)), dist_saved, hidden_states, attention_mask, context, context_mask, rotary_pos_emb)
                i += k
        elif self.config.recompute_method == "block":
            for i in range(n):
                if i < k:
                    hidden_states, context = checkpoint(custom(i, i + 1), dist_saved, hidden_states, attention_mask, context, context_mask, rotary_pos_emb)
                else:
                    hidden_states, context = custom(i, i + 1)(hidden_states, attention_mask, context, context_mask, rotary_pos_emb)
        else:
            raise ValueError("invalid activation recompute method")
        return hidden_states

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None, rotary_pos_emb=None,
                rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None, inference_context=None, packed_seq_params=None,
                sequence_len_offset=None, *, inference_params=None, **kwargs):
        inference_context = inference_context or inference_params
        if not self.pre_process:
            hidden_states = self.input_tensor
        hidden_states = make_viewless_tensor(hidden_states, requires_grad=True, keep_graph=True)
        rng_ctx = get_cuda_rng_tracker().fork() if (self.config.sequence_parallel and get_cuda_rng_tracker().is_initialized()) else nullcontext()
        mhc = getattr(self.config, "enable_mhc_connections", False)
        if mhc:
            from .hyper_connection import HyperConnectionModule

            hidden_states = HyperConnectionModule.input_expand(hidden_states, self.config.mhc_num_residual_streams)      # [s, b, C] → n streams
        with rng_ctx:
            if self.config.recompute_granularity == "full" and self.training:
                hidden_states = self._checkpointed_forward(hidden_states, attention_mask, context, context_mask, rotary_pos_emb, attention_bias, packed_seq_params)
            else:
                from ..fp8_utils import get_fp8_context

                for layer in self.layers:
                    # per-layer FP8 scope: the linear layers inside consult fp8_enabled() (first / last layers may stay bf16)
                    with get_fp8_context(self.config, layer.layer_number - 1):
                        hidden_states, context = layer(
                            hidden_states, attention_mask=attention_mask, context=context, context_mask=context_mask,
                            rotary_pos_emb=rotary_pos_emb, attention_bias=attention_bias, inference_context=inference_context,
                            packed_seq_params=packed_seq_params,
                        )
        if mhc:
            hidden_states = HyperConnectionModule.output_contract(hidden_states, self.config.mhc_num_residual_streams)   # mean of the streams
        if self.final_layernorm is not None:
            hidden_states = self.final_layernorm(hidden_states)
            hidden_states = make_viewless_tensor(hidden_states, requires_grad=True, keep_graph=True)
        return hidden_states

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: dict = None) -> ShardedStateDict:
        """Homogeneous stacks are stored under ONE key per parameter with a prepended
        layer axis (offset = global layer index), so PP/VPP re-partitioning is free
        (reference :774-845)."""
        assert not sharded_offsets, "unexpected sharded offsets"
        non_homogeneous = (metadata or {}).get("non_homogeneous_layers", False) or self.config.num_moe_experts is not None or self.config.heterogeneous_block_specs
        out = {}
        layer_prefix = f"{prefix}layers."
        n_total = self.config.num_layers
        for layer in self.layers:
            off = layer.layer_number - 1  # global index (TransformerLayer added the pp offset)
            local_prefix = f"{layer_prefix}{layer.layer_number - 1 - get_transformer_layer_offset(self.config, self.vp_stage)}."
            if non_homogeneous:
                sub_prefix, so = f"{layer_prefix}{off}.", []
            else:
                sub_prefix, so = layer_prefix, [(0, off, n_total)]
            sd = layer.sharded_state_dict(local_prefix, so, metadata)
            # re-key: local module path → global key
            for k, v in sd.items():
                if hasattr(v, "key"):
                    v.key = v.key.replace(local_prefix, sub_prefix, 1)
                out[k] = v
        for name, module in self.named_children():
            if module is not self.layers:
                out.update(sharded_state_dict_default(module, f"{prefix}{name}.", sharded_offsets, metadata))
        return out
