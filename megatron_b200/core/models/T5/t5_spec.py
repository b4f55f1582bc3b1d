"""T5 encoder / decoder layer specs (reference ``models/T5/t5_spec.py``)."""
from ...transformer.attention import CrossAttention, CrossAttentionSubmodules, SelfAttention, SelfAttentionSubmodules
from ...transformer.enums import AttnMaskType
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlockSubmodules
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add
from ..backends import B200SpecProvider


def _mlp(b):
    return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=b.column_parallel_linear(), linear_fc2=b.row_parallel_linear()))


def encoder_model_with_local_spec() -> ModuleSpec:
    b = B200SpecProvider()
    return ModuleSpec(
        module=TransformerLayer,
        submodules=TransformerLayerSubmodules(
            input_layernorm=b.layer_norm(),
            self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": AttnMaskType.padding},
                                      submodules=SelfAttentionSubmodules(linear_qkv=b.column_parallel_linear(), core_attention=b.core_attention(), linear_proj=b.row_parallel_linear())),
            self_attn_bda=get_bias_dropout_add, pre_mlp_layernorm=b.layer_norm(), mlp=_mlp(b), mlp_bda=get_bias_dropout_add,
        ),
    )


def decoder_model_with_local_spec() -> ModuleSpec:
    b = B200SpecProvider()
    return ModuleSpec(
        module=TransformerLayer,
        submodules=TransformerLayerSubmodules(
            input_layernorm=b.layer_norm(),
            self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": AttnMaskType.causal},
                                      submodules=SelfAttentionSubmodules(linear_qkv=b.column_parallel_linear(), core_attention=b.core_attention(), linear_proj=b.row_parallel_linear())),
            self_attn_bda=get_bias_dropout_add,
            pre_cross_attn_layernorm=b.layer_norm(),
            cross_attention=ModuleSpec(module=CrossAttention, params={"attn_mask_type": AttnMaskType.arbitrary},
                                       submodules=CrossAttentionSubmodules(linear_q=b.column_parallel_linear(), linear_kv=b.column_parallel_linear(),
                                                                           core_attention=b.core_attention(), linear_proj=b.row_parallel_linear())),
            cross_attn_bda=get_bias_dropout_add,
            pre_mlp_layernorm=b.layer_norm(), mlp=_mlp(b), mlp_bda=get_bias_dropout_add,
        ),
    )


def get_t5_encoder_with_local_block_spec(num_layers: int) -> TransformerBlockSubmodules:
    from ...transformer.torch_norm import FusedNorm

    return TransformerBlockSubmodules(layer_specs=[encoder_model_with_local_spec()] * num_layers, layer_norm=FusedNorm)


def get_t5_decoder_with_local_block_spec(num_layers: int) -> TransformerBlockSubmodules:
    from ...transformer.torch_norm import FusedNorm

    return TransformerBlockSubmodules(layer_specs=[decoder_model_with_local_spec()] * num_layers, layer_norm=FusedNorm)


get_t5_encoder_with_transformer_engine_block_spec = get_t5_encoder_with_local_block_spec
get_t5_decoder_with_transformer_engine_block_spec = get_t5_decoder_with_local_block_spec
