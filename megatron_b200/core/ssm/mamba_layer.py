"""Pre-norm residual wrapper around a mixer (reference ``ssm/mamba_layer.py``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import torch

from ..transformer.identity_op import IdentityOp
from ..transformer.module import GraphableMegatronModule
from ..transformer.spec_utils import ModuleSpec, build_module
from ..transformer.transformer_config import TransformerConfig


@dataclass
class MambaLayerSubmodules:
    norm: Union[ModuleSpec, type] = IdentityOp
    mixer: Union[ModuleSpec, type] = IdentityOp
    mamba_bda: Union[ModuleSpec, type] = IdentityOp


class MambaLayer(GraphableMegatronModule):
    def __init__(self, config: TransformerConfig, submodules: MambaLayerSubmodules, layer_number: int = 1, residual_in_fp32: bool = False,
                 pg_collection=None, **_):
        super().__init__(config)
        self.layer_number = layer_number
        self.residual_in_fp32 = residual_in_fp32
        self.hidden_dropout = config.hidden_dropout
        self.norm = build_module(submodules.norm, config=config, hidden_size=config.hidden_size, eps=config.layernorm_epsilon)
        self.mixer = build_module(submodules.mixer, config, d_model=config.hidden_size, layer_number=layer_number, pg_collection=pg_collection)
        self.mamba_bda = build_module(submodules.mamba_bda)

    def forward(self, hidden_states, attention_mask=None, inference_context=None, rotary_pos_emb=None, *, inference_params=None, **_):
        residual = hidden_states.float() if self.residual_in_fp32 else hidden_states
        h = self.norm(hidden_states.to(dtype=self.config.params_dtype))
        out_with_bias = self.mixer(h, inference_context=inference_context or inference_params)
        with torch.enable_grad():
            return self.mamba_bda(self.training, self.config.bias_dropout_fusion)(out_with_bias, residual, self.hidden_dropout)
