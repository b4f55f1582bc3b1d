"""Specs for the hybrid Mamba stack (reference ``models/mamba/mamba_layer_specs.py``)."""
from ...ssm.gated_delta_net import GatedDeltaNet, GatedDeltaNetSubmodules
from ...ssm.mamba_block import MambaStack, MambaStackSubmodules
from ...ssm.mamba_layer import MambaLayer, MambaLayerSubmodules
from ...ssm.mamba_mixer import MambaMixer, MambaMixerSubmodules
from ...transformer.attention import SelfAttention, SelfAttentionSubmodules
from ...transformer.enums import AttnMaskType
from ...transformer.experimental_attention_variant import DSAMLASelfAttention
from ...transformer.identity_op import IdentityOp
from ...transformer.multi_latent_attention import MLASelfAttention, MLASelfAttentionSubmodules
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add
from ..backends import B200SpecProvider

_b = B200SpecProvider()


def _latent_attention_layer(module):
    """Attention-only layer around multi-latent attention (``+``) or its sparse variant (``D``); needs an ``MLATransformerConfig``."""
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=_b.layer_norm(),
        self_attention=ModuleSpec(module=module, params={"attn_mask_type": AttnMaskType.causal}, submodules=MLASelfAttentionSubmodules(
            linear_q_proj=_b.column_parallel_linear(), linear_q_down_proj=_b.linear(), linear_q_up_proj=_b.column_parallel_linear(), linear_kv_down_proj=_b.linear(),
            linear_kv_up_proj=_b.column_parallel_linear(), core_attention=_b.core_attention(), linear_proj=_b.row_parallel_linear(), q_layernorm=IdentityOp,
            kv_layernorm=IdentityOp)),
        self_attn_bda=get_bias_dropout_add))


mamba_stack_spec = ModuleSpec(
    module=MambaStack,
    submodules=MambaStackSubmodules(
        mamba_layer=ModuleSpec(
            module=MambaLayer,
            submodules=MambaLayerSubmodules(
                norm=_b.layer_norm(),
                mixer=ModuleSpec(module=MambaMixer, submodules=MambaMixerSubmodules(in_proj=_b.column_parallel_linear(), out_proj=_b.row_parallel_linear())),
                mamba_bda=get_bias_dropout_add,
            ),
        ),
        gdn_layer=ModuleSpec(                   # ``G`` in the hybrid pattern: the same pre-norm residual wrapper around a gated-delta-net mixer
            module=MambaLayer,
            submodules=MambaLayerSubmodules(
                norm=_b.layer_norm(),
                mixer=ModuleSpec(module=GatedDeltaNet, submodules=GatedDeltaNetSubmodules(in_proj=_b.column_parallel_linear(), out_proj=_b.row_parallel_linear())),
                mamba_bda=get_bias_dropout_add,
            ),
        ),
        mla_layer=_latent_attention_layer(MLASelfAttention),
        dsa_layer=_latent_attention_layer(DSAMLASelfAttention),
        attention_layer=ModuleSpec(
            module=TransformerLayer,
            submodules=TransformerLayerSubmodules(
                input_layernorm=_b.layer_norm(),
                self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": AttnMaskType.causal},
                                          submodules=SelfAttentionSubmodules(linear_qkv=_b.column_parallel_linear(), core_attention=_b.core_attention(),
                                                                             linear_proj=_b.row_parallel_linear())),
                self_attn_bda=get_bias_dropout_add,
            ),
        ),
        mlp_layer=ModuleSpec(
            module=TransformerLayer,
            submodules=TransformerLayerSubmodules(
                pre_mlp_layernorm=_b.layer_norm(),
                mlp=ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=_b.column_parallel_linear(), linear_fc2=_b.row_parallel_linear())),
                mlp_bda=get_bias_dropout_add,
            ),
        ),
    ),
)
