"""Layer specs for GPT-family models (reference ``models/gpt/gpt_layer_specs.py:351-799``)."""
from __future__ import annotations

from typing import Optional

from ...transformer.attention import SelfAttention, SelfAttentionSubmodules
from ...transformer.enums import AttnMaskType
from ...transformer.identity_op import IdentityOp
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.spec_utils import ModuleSpec
from ...transformer.torch_norm import L2Norm
from ...transformer.transformer_block import TransformerBlockSubmodules, get_num_layers_to_build
from ...transformer.transformer_config import TransformerConfig
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add, get_transformer_layer_offset
from ..backends import B200SpecProvider, BackendSpecProvider


def get_mlp_module_spec_for_backend(backend: BackendSpecProvider, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                                    moe_use_legacy_grouped_gemm: bool = False, use_te_activation_func: bool = False) -> ModuleSpec:
    if num_experts is None:
        return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=backend.column_parallel_linear(), linear_fc2=backend.row_parallel_linear()))
    from .moe_module_specs import get_moe_module_spec_for_backend

    return get_moe_module_spec_for_backend(backend, num_experts, moe_grouped_gemm, moe_use_legacy_grouped_gemm)


def get_mlp_module_spec(use_te: bool = False, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False, fp8=None,
                        moe_use_legacy_grouped_gemm: bool = False) -> ModuleSpec:
    return get_mlp_module_spec_for_backend(B200SpecProvider(), num_experts, moe_grouped_gemm, moe_use_legacy_grouped_gemm)


def get_gpt_layer_local_spec(num_experts: Optional[int] = None, moe_grouped_gemm: bool = False, qk_layernorm: bool = False,
                             multi_latent_attention: bool = False, fp8: Optional[str] = None, moe_use_legacy_grouped_gemm: bool = False,
                             normalization: Optional[str] = None, qk_l2_norm: bool = False, use_kitchen: bool = False,
                             use_te_activation_func: bool = False, **kwargs) -> ModuleSpec:
    """Spec of one decoder layer.  Names of the norm params are remapped on checkpoint to the
    fused names (``linear_qkv.layer_norm_weight``) so checkpoints interoperate with the
    reference's TE layout (reference :459-462)."""
    backend = B200SpecProvider()
    norm = backend.layer_norm(rms_norm=normalization == "RMSNorm")
    qk_norm = L2Norm if qk_l2_norm else (backend.layer_norm(for_qk=True) if qk_layernorm else None)
    mlp = get_mlp_module_spec_for_backend(backend, num_experts, moe_grouped_gemm, moe_use_legacy_grouped_gemm)
    if multi_latent_attention:
        from ...transformer.multi_latent_attention import MLASelfAttention, MLASelfAttentionSubmodules

        attn = ModuleSpec(
            module=MLASelfAttention, params={"attn_mask_type": AttnMaskType.causal},
            submodules=MLASelfAttentionSubmodules(
                linear_q_proj=backend.column_parallel_linear(), linear_q_down_proj=backend.linear(),
                linear_q_up_proj=backend.column_parallel_linear(), linear_kv_down_proj=backend.linear(),
                linear_kv_up_proj=backend.column_parallel_linear(), core_attention=backend.core_attention(),
                linear_proj=backend.row_parallel_linear(), q_layernorm=norm if qk_layernorm else IdentityOp,
                kv_layernorm=norm if qk_layernorm else IdentityOp,
            ),
        )
        keys_map = {}
    else:
        attn = ModuleSpec(
            module=SelfAttention, params={"attn_mask_type": AttnMaskType.causal},
            submodules=SelfAttentionSubmodules(
                linear_qkv=backend.column_parallel_linear(), core_attention=backend.core_attention(),
                linear_proj=backend.row_parallel_linear(), q_layernorm=qk_norm, k_layernorm=qk_norm,
            ),
        )
        keys_map = {"input_layernorm.": "self_attention.linear_qkv.layer_norm_", "pre_mlp_layernorm.": "mlp.linear_fc1.layer_norm_"}
    if num_experts is not None:
        keys_map = {k: v for k, v in keys_map.items() if not k.startswith("pre_mlp")}
    return ModuleSpec(
        module=TransformerLayer,
        submodules=TransformerLayerSubmodules(
            input_layernorm=norm, self_attention=attn, self_attn_bda=get_bias_dropout_add,
            pre_mlp_layernorm=norm, mlp=mlp, mlp_bda=get_bias_dropout_add, sharded_state_dict_keys_map=keys_map,
        ),
    )


# this framework has a single native backend; the TE-named entry point resolves to it
get_gpt_layer_with_transformer_engine_spec = get_gpt_layer_local_spec
get_gpt_layer_b200_spec = get_gpt_layer_local_spec


def get_gpt_decoder_block_spec(config: TransformerConfig, use_transformer_engine: bool = False, normalization: Optional[str] = None,
                               qk_l2_norm: bool = False, vp_stage: Optional[int] = None, pp_rank: Optional[int] = None) -> TransformerBlockSubmodules:
    """Per-layer specs for models that interleave dense and MoE layers (``moe_layer_freq``)."""
    dense = get_gpt_layer_local_spec(None, False, config.qk_layernorm, config.multi_latent_attention, normalization=normalization or config.normalization, qk_l2_norm=qk_l2_norm)
    moe = get_gpt_layer_local_spec(config.num_moe_experts, config.moe_grouped_gemm, config.qk_layernorm, config.multi_latent_attention,
                                   normalization=normalization or config.normalization, qk_l2_norm=qk_l2_norm)
    freq = config.moe_layer_freq
    if isinstance(freq, int):
        pattern = [1 if (i % freq == 0) else 0 for i in range(config.num_layers)]
    else:
        pattern = list(freq)
        assert len(pattern) == config.num_layers, "moe_layer_freq pattern length must equal num_layers"
    if config.num_moe_experts is None:
        pattern = [0] * config.num_layers
    specs = [moe if p else dense for p in pattern]
    n = get_num_layers_to_build(config, vp_stage, pp_rank)
    off = get_transformer_layer_offset(config, vp_stage, pp_rank)
    from ...transformer.torch_norm import FusedNorm

    return TransformerBlockSubmodules(layer_specs=specs[off : off + n], layer_norm=FusedNorm)


def get_gpt_mtp_block_spec(config: TransformerConfig, spec, use_transformer_engine: bool = False, vp_stage: Optional[int] = None):
    """MTP block spec built from a decoder-layer spec (reference ``gpt_layer_specs.py:get_gpt_mtp_block_spec``)."""
    from ...transformer.multi_token_prediction import get_mtp_block_spec

    layer_spec = spec.layer_specs[-1] if isinstance(spec, TransformerBlockSubmodules) else spec
    return get_mtp_block_spec(config, layer_spec, use_transformer_engine, vp_stage)
