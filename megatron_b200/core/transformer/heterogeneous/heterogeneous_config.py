"""Heterogeneous (per-layer different) transformer architectures — Nemotron-NAS style (reference ``transformer/heterogeneous/`` 379 LoC +
``models/gpt/heterogeneous/heterogeneous_layer_specs.py``).

``heterogeneous_layers_config_json`` (HF ``block_configs`` format) lists, per layer, how its attention and MLP differ from the base config:

```json
{"block_configs": [
  {"attention": {"no_op": false, "n_heads_in_group": 4},     "ffn": {"no_op": false, "ffn_mult": 2.625}},
  {"attention": {"no_op": true},                              "ffn": {"no_op": false, "ffn_mult": 1.3125}},
  {"attention": {"replace_with_linear": true},                "ffn": {"no_op": true}}
]}
```
``no_op`` removes the sub-block (and its norm); ``replace_with_linear`` swaps it for a single hidden→hidden linear; ``n_heads_in_group`` sets the
GQA group size; ``ffn_mult`` the FFN width (rounded up to a multiple of 256 like the reference)."""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from ..attention import SelfAttention, SelfAttentionSubmodules
from ..enums import AttnMaskType
from ..identity_op import IdentityFuncOp, IdentityOp
from ..mlp import MLP, MLPSubmodules
from ..spec_utils import ModuleSpec
from ..transformer_block import TransformerBlockSubmodules
from ..transformer_config import TransformerConfig
from ..transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add


@dataclass
class AttentionBlockConfig:
    no_op: bool = False
    replace_with_linear: bool = False
    num_query_groups: Optional[int] = None


@dataclass
class MLPBlockConfig:
    no_op: bool = False
    replace_with_linear: bool = False
    ffn_hidden_size: Optional[int] = None


@dataclass
class BlockConfig:
    attention: AttentionBlockConfig = field(default_factory=AttentionBlockConfig)
    mlp: MLPBlockConfig = field(default_factory=MLPBlockConfig)


def _ffn_mult_to_size(mult: float, hidden: int, multiple_of: int = 256) -> int:
    size = int(2 * mult * hidden / 3)
    return multiple_of * ((size + multiple_of - 1) // multiple_of)


@dataclass
class HeterogeneousTransformerConfig(TransformerConfig):
    heterogeneous_layers_config_path: Optional[str] = None
    heterogeneous_layers_config_encoded_json: Optional[str] = None
    per_block_parameters: List[BlockConfig] = field(default_factory=list)

    def __post_init__(self):
        super().__post_init__()
        raw = self.heterogeneous_layers_config_encoded_json
        if raw is None and self.heterogeneous_layers_config_path:
            with open(self.heterogeneous_layers_config_path) as f:
                raw = f.read()
        if raw is None:
            return
        blocks = json.loads(raw)["block_configs"]
        assert len(blocks) == self.num_layers, f"{len(blocks)} block configs for {self.num_layers} layers"
        out = []
        for b in blocks:
            a, m = b.get("attention", {}) or {}, b.get("ffn", b.get("mlp", {})) or {}
            groups = None
            if a.get("n_heads_in_group"):
                groups = self.num_attention_heads // int(a["n_heads_in_group"])
            elif a.get("num_query_groups"):
                groups = int(a["num_query_groups"])
            ffn = None
            if m.get("ffn_mult") is not None:
                ffn = _ffn_mult_to_size(float(m["ffn_mult"]), self.hidden_size)
            elif m.get("ffn_hidden_size"):
                ffn = int(m["ffn_hidden_size"])
            out.append(BlockConfig(AttentionBlockConfig(bool(a.get("no_op")), bool(a.get("replace_with_linear")), groups),
                                   MLPBlockConfig(bool(m.get("no_op")), bool(m.get("replace_with_linear")), ffn)))
        self.per_block_parameters = out


class _LinearSubBlock(torch.nn.Module):
    """hidden → hidden linear standing in for a whole attention or MLP sub-block (``replace_with_linear``)."""

    def __init__(self, config, layer_number: int = 1, **_):
        super().__init__()
        from ...tensor_parallel.layers import ColumnParallelLinear

        self.proj = ColumnParallelLinear(config.hidden_size, config.hidden_size, config=config, init_method=config.init_method, bias=False, gather_output=True,
                                         skip_bias_add=True)

    def forward(self, hidden_states, *args, **kwargs):
        return self.proj(hidden_states)


class _GroupsOverrideAttention(SelfAttention):
    """SelfAttention with a per-layer number of query groups."""

    def __init__(self, config, submodules, layer_number, num_query_groups: Optional[int] = None, **kw):
        if num_query_groups is not None:
            import copy

            config = copy.copy(config)
            config.num_query_groups = num_query_groups
        super().__init__(config, submodules, layer_number, **kw)


def get_gpt_heterogeneous_layer_spec(config: HeterogeneousTransformerConfig, use_te: bool = False) -> TransformerBlockSubmodules:
    from ...models.backends import B200SpecProvider
    from ..torch_norm import FusedNorm

    b = B200SpecProvider()
    norm = b.layer_norm()
    specs = []
    for blk in config.per_block_parameters or [BlockConfig() for _ in range(config.num_layers)]:
        a, m = blk.attention, blk.mlp
        if a.no_op:
            attn, attn_norm, attn_bda = IdentityOp, IdentityOp, IdentityFuncOp
        elif a.replace_with_linear:
            attn, attn_norm, attn_bda = ModuleSpec(module=_LinearSubBlock), norm, get_bias_dropout_add
        else:
            attn = ModuleSpec(module=_GroupsOverrideAttention, params={"attn_mask_type": AttnMaskType.causal, "num_query_groups": a.num_query_groups},
                              submodules=SelfAttentionSubmodules(linear_qkv=b.column_parallel_linear(), core_attention=b.core_attention(), linear_proj=b.row_parallel_linear()))
            attn_norm, attn_bda = norm, get_bias_dropout_add
        if m.no_op:
            mlp, mlp_norm, mlp_bda = IdentityOp, IdentityOp, IdentityFuncOp
        elif m.replace_with_linear:
            mlp, mlp_norm, mlp_bda = ModuleSpec(module=_LinearSubBlock), norm, get_bias_dropout_add
        else:
            mlp = ModuleSpec(module=MLP, params={"ffn_hidden_size": m.ffn_hidden_size} if m.ffn_hidden_size else {},
                             submodules=MLPSubmodules(linear_fc1=b.column_parallel_linear(), linear_fc2=b.row_parallel_linear()))
            mlp_norm, mlp_bda = norm, get_bias_dropout_add
        specs.append(ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
            input_layernorm=attn_norm, self_attention=attn, self_attn_bda=attn_bda, pre_mlp_layernorm=mlp_norm, mlp=mlp, mlp_bda=mlp_bda)))
    return TransformerBlockSubmodules(layer_specs=specs, layer_norm=FusedNorm)
