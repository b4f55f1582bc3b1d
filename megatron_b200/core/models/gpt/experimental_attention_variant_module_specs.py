"""Layer specs for the experimental attention variants (reference ``models/gpt/experimental_attention_variant_module_specs.py``).

``config.experimental_attention_variant``:

* ``"gdn"`` (alias ``"gated_delta_net"``) — linear attention: gated-delta-net mixers replace softmax attention in the layers that
  ``linear_attention_freq`` marks (an int N = one softmax layer after every N-1 linear ones, or an explicit 0/1 list);
* ``"dsa"`` — every attention layer is absorbed MLA with the DeepSeek sparse-attention core (needs an ``MLATransformerConfig``).

``get_transformer_block_with_experimental_attention_variant_spec`` returns the per-layer spec list of THIS pipeline stage, MoE pattern included."""
from __future__ import annotations

import warnings
from typing import List, Optional

from ...transformer.enums import AttnMaskType
from ...transformer.identity_op import IdentityOp
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlockSubmodules
from ...transformer.transformer_config import TransformerConfig
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add
from ..backends import B200SpecProvider
from .gpt_layer_specs import get_mlp_module_spec_for_backend


def normalize_experimental_attention_variant(name: Optional[str]) -> Optional[str]:
    if name == "gated_delta_net":
        warnings.warn("experimental_attention_variant='gated_delta_net' is deprecated: use 'gdn'", DeprecationWarning)
        return "gdn"
    return name


def is_gated_delta_net_variant(name: Optional[str]) -> bool:
    return normalize_experimental_attention_variant(name) in ("gdn", "gdn2") if name else False


def is_linear_attention_variant(name: Optional[str]) -> bool:
    return is_gated_delta_net_variant(name)


def get_linear_attention_pattern(config: TransformerConfig) -> List[int]:
    """1 = linear-attention layer, 0 = softmax-attention layer."""
    f = getattr(config, "linear_attention_freq", None)
    if isinstance(f, int):
        return [0 if (i + 1) % f == 0 else 1 for i in range(config.num_layers)]
    if isinstance(f, (list, tuple)):
        if len(f) != config.num_layers:
            raise ValueError(f"linear_attention_freq lists {len(f)} layers, the model has {config.num_layers}")
        return [int(x) for x in f]
    if is_linear_attention_variant(getattr(config, "experimental_attention_variant", None)):
        raise ValueError("linear_attention_freq must be set for a linear-attention variant")
    return [0] * config.num_layers


def get_moe_layer_pattern(config: TransformerConfig) -> List[int]:
    """1 = MoE layer, 0 = dense MLP (``moe_layer_freq``: int N = every N-th layer, or an explicit list)."""
    if config.num_moe_experts is None:
        return [0] * config.num_layers
    f = getattr(config, "moe_layer_freq", 1)
    if isinstance(f, int):
        return [1 if i % f == 0 else 0 for i in range(config.num_layers)]
    if len(f) != config.num_layers:
        raise ValueError(f"moe_layer_freq lists {len(f)} layers, the model has {config.num_layers}")
    return [int(x) for x in f]


def get_gated_delta_net_module_spec(backend: Optional[B200SpecProvider] = None) -> ModuleSpec:
    from ...ssm.gated_delta_net import GatedDeltaNet, GatedDeltaNetSubmodules

    b = backend or B200SpecProvider()
    return ModuleSpec(module=GatedDeltaNet, submodules=GatedDeltaNetSubmodules(in_proj=b.column_parallel_linear(), out_proj=b.row_parallel_linear()))


def get_dsa_module_spec_for_backend(config: TransformerConfig, backend: Optional[B200SpecProvider] = None) -> ModuleSpec:
    from ...transformer.experimental_attention_variant import DSAMLASelfAttention
    from ...transformer.multi_latent_attention import MLASelfAttentionSubmodules

    if not getattr(config, "multi_latent_attention", False):
        raise ValueError("the dsa variant runs on the MLA latent: it needs multi_latent_attention (MLATransformerConfig)")
    b = backend or B200SpecProvider()
    norm = b.layer_norm(rms_norm=config.normalization == "RMSNorm")
    return ModuleSpec(module=DSAMLASelfAttention, params={"attn_mask_type": AttnMaskType.causal}, submodules=MLASelfAttentionSubmodules(
        linear_q_proj=b.column_parallel_linear(), linear_q_down_proj=b.linear(), linear_q_up_proj=b.column_parallel_linear(), linear_kv_down_proj=b.linear(),
        linear_kv_up_proj=b.column_parallel_linear(), core_attention=b.core_attention(), linear_proj=b.row_parallel_linear(),
        q_layernorm=norm if config.qk_layernorm else IdentityOp, kv_layernorm=norm if config.qk_layernorm else IdentityOp))


def get_experimental_attention_variant_module_spec(config: TransformerConfig, backend: Optional[B200SpecProvider] = None) -> ModuleSpec:
    v = normalize_experimental_attention_variant(config.experimental_attention_variant)
    if is_gated_delta_net_variant(v):
        return get_gated_delta_net_module_spec(backend)
    if v == "dsa":
        return get_dsa_module_spec_for_backend(config, backend)
    raise ValueError(f"unknown experimental_attention_variant {v!r}")


_ADAPTER = None


def _gdn_attention_adapter():
    """A gated-delta-net mixer with the call signature of a ``self_attention`` module, so it can sit in that slot of a transformer layer."""
    global _ADAPTER
    if _ADAPTER is None:
        from ...ssm.gated_delta_net import GatedDeltaNet

        class GatedDeltaNetAttention(GatedDeltaNet):
            def __init__(self, config, submodules, layer_number: int = 1, attn_mask_type=None, cp_comm_type=None, pg_collection=None, **_):
                super().__init__(config, submodules, layer_number=layer_number, pg_collection=pg_collection)

            def forward(self, hidden_states, attention_mask=None, inference_context=None, *, inference_params=None, **_):
                return super().forward(hidden_states, inference_context=inference_context or inference_params)

        _ADAPTER = GatedDeltaNetAttention
    return _ADAPTER


def get_transformer_layer_with_experimental_attention_variant_spec(config: TransformerConfig, linear: bool, moe: bool, backend: Optional[B200SpecProvider] = None) -> ModuleSpec:
    """One decoder layer: its attention slot holds the variant (or plain / MLA attention when ``linear`` is False under the gdn variant)."""
    from .gpt_layer_specs import get_gpt_layer_local_spec

    b = backend or B200SpecProvider()
    v = normalize_experimental_attention_variant(config.experimental_attention_variant)
    base = get_gpt_layer_local_spec(config.num_moe_experts if moe else None, config.moe_grouped_gemm, config.qk_layernorm, getattr(config, "multi_latent_attention", False),
                                    normalization=config.normalization)
    if v == "dsa":
        attn = get_dsa_module_spec_for_backend(config, b)
    elif linear:
        gdn = get_gated_delta_net_module_spec(b)
        attn = ModuleSpec(module=_gdn_attention_adapter(), submodules=gdn.submodules)
    else:
        return base
    sub = base.submodules
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=sub.input_layernorm, self_attention=attn, self_attn_bda=get_bias_dropout_add, pre_mlp_layernorm=sub.pre_mlp_layernorm, mlp=sub.mlp,
        mlp_bda=get_bias_dropout_add))


def get_transformer_block_with_experimental_attention_variant_spec(config: TransformerConfig, vp_stage: Optional[int] = None) -> TransformerBlockSubmodules:
    """Per-layer specs of the layers this pipeline (and virtual) stage owns."""
    from ...transformer.transformer_block import get_num_layers_to_build
    from ...transformer.transformer_layer import get_transformer_layer_offset

    b = B200SpecProvider()
    la, moe = get_linear_attention_pattern(config), get_moe_layer_pattern(config)
    specs = [get_transformer_layer_with_experimental_attention_variant_spec(config, bool(la[i]), bool(moe[i]), b) for i in range(config.num_layers)]
    off, n = get_transformer_layer_offset(config, vp_stage=vp_stage), get_num_layers_to_build(config, vp_stage=vp_stage)
    return TransformerBlockSubmodules(layer_specs=specs[off : off + n], layer_norm=b.layer_norm(rms_norm=config.normalization == "RMSNorm"))


__all__ = ["get_dsa_module_spec_for_backend", "get_experimental_attention_variant_module_spec", "get_gated_delta_net_module_spec", "get_linear_attention_pattern",
           "get_moe_layer_pattern", "get_transformer_block_with_experimental_attention_variant_spec", "get_transformer_layer_with_experimental_attention_variant_spec",
           "is_gated_delta_net_variant", "is_linear_attention_variant", "normalize_experimental_attention_variant", "get_mlp_module_spec_for_backend"]
