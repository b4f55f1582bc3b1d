"""Stack of Mamba / attention / MLP layers following a hybrid pattern (reference ``ssm/mamba_block.py`` ``MambaStack``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch

from .. import parallel_state as ps
from ..transformer.identity_op import IdentityOp
from ..transformer.module import MegatronModule
from ..transformer.spec_utils import ModuleSpec, build_module
from ..transformer.transformer_config import TransformerConfig
from ..utils import make_viewless_tensor
from .mamba_hybrid_layer_allocation import Symbols, allocate_layers, parse_hybrid_pattern


@dataclass
class MambaStackSubmodules:
    mamba_layer: Union[ModuleSpec, type] = IdentityOp
    attention_layer: Union[ModuleSpec, type] = IdentityOp
    mlp_layer: Union[ModuleSpec, type] = IdentityOp
    moe_layer: Union[ModuleSpec, type] = IdentityOp
    gdn_layer: Union[ModuleSpec, type] = IdentityOp
    mla_layer: Union[ModuleSpec, type] = IdentityOp
    dsa_layer: Union[ModuleSpec, type] = IdentityOp


class MambaStack(MegatronModule):
    def __init__(self, config: TransformerConfig, submodules: MambaStackSubmodules, residual_in_fp32: bool = False, pre_process: bool = True,
                 hybrid_attention_ratio: float = 0.0, hybrid_mlp_ratio: float = 0.0, hybrid_override_pattern: Optional[str] = None,
                 post_layer_norm: bool = True, post_process: bool = True, device=None, dtype=None, pg_collection=None, vp_stage=None):
        super().__init__(config)
        self.pre_process, self.post_process, self.post_layer_norm = pre_process, post_process, post_layer_norm
        self.input_tensor = None
        layout = allocate_layers(config.num_layers, hybrid_attention_ratio, hybrid_mlp_ratio, hybrid_override_pattern)
        pp = ps.get_pipeline_model_parallel_world_size() if ps.model_parallel_is_initialized() else 1
        segments = parse_hybrid_pattern(hybrid_override_pattern)[1] if hybrid_override_pattern else None
        if segments is not None and pp > 1:
            # "|" in the pattern: explicit (possibly uneven) pipeline split
            if len(segments) != pp:
                raise ValueError(f"the hybrid pattern has {len(segments)} pipeline segments, the job has {pp} pipeline stages")
            r = ps.get_pipeline_model_parallel_rank()
            off, per = sum(segments[:r]), segments[r]
        else:
            if config.num_layers % pp != 0:
                raise ValueError(f"{config.num_layers} layers cannot be split evenly over {pp} pipeline stages: mark the stage boundaries with '|' in the hybrid pattern")
            per = config.num_layers // pp
            off = (ps.get_pipeline_model_parallel_rank() if pp > 1 else 0) * per
        self.layer_type_list = layout[off : off + per]
        self.layers = torch.nn.ModuleList()
        kw = {"pg_collection": pg_collection} if pg_collection is not None else {}
        for i, sym in enumerate(self.layer_type_list):
            n = off + i + 1
            if sym == Symbols.MAMBA:
                layer = build_module(submodules.mamba_layer, config=config, residual_in_fp32=residual_in_fp32, layer_number=n, **kw)
            elif sym == Symbols.GDN:
                layer = build_module(submodules.gdn_layer, config=config, residual_in_fp32=residual_in_fp32, layer_number=n, **kw)
            elif sym == Symbols.ATTENTION:
                layer = build_module(submodules.attention_layer, config=config, layer_number=n, **kw)
            elif sym == Symbols.MLA:
                layer = build_module(submodules.mla_layer, config=config, layer_number=n, **kw)
            elif sym == Symbols.DS_ATTENTION:
                layer = build_module(submodules.dsa_layer, config=config, layer_number=n, **kw)
            elif sym == Symbols.MLP:
                layer = build_module(submodules.mlp_layer, config=config, layer_number=n, **kw)
            else:
                layer = build_module(submodules.moe_layer, config=config, layer_number=n, **kw)
            self.layers.append(layer)
        if self.post_process and self.post_layer_norm:
            from ..transformer.torch_norm import FusedNorm

            self.final_norm = FusedNorm(config, config.hidden_size, eps=config.layernorm_epsilon)

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor

    def forward(self, hidden_states, attention_mask=None, inference_context=None, rotary_pos_emb=None, *, inference_params=None, **kw):
        if not self.pre_process:
            hidden_states = self.input_tensor
        inference_context = inference_context or inference_params
        hidden_states = make_viewless_tensor(hidden_states, requires_grad=True, keep_graph=True)
        for layer in self.layers:
            out = layer(hidden_states, attention_mask=attention_mask, inference_context=inference_context, rotary_pos_emb=rotary_pos_emb)
            hidden_states = out[0] if isinstance(out, tuple) else out
        if self.post_process and self.post_layer_norm:
            hidden_states = self.final_norm(hidden_states)
        return hidden_states
