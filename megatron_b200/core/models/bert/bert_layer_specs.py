"""BERT encoder layer spec: bidirectional self-attention with a padding mask (reference ``models/bert/bert_layer_specs.py``)."""
from ...transformer.attention import SelfAttention, SelfAttentionSubmodules
from ...transformer.enums import AttnMaskType
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_bias_dropout_add
from ..backends import B200SpecProvider


def get_bert_layer_local_spec() -> ModuleSpec:
    b = B200SpecProvider()
    return ModuleSpec(
        module=TransformerLayer,
        submodules=TransformerLayerSubmodules(
            input_layernorm=b.layer_norm(),
            self_attention=ModuleSpec(
                module=SelfAttention, params={"attn_mask_type": AttnMaskType.padding},
                submodules=SelfAttentionSubmodules(linear_qkv=b.column_parallel_linear(), core_attention=b.core_attention(), linear_proj=b.row_parallel_linear()),
            ),
            self_attn_bda=get_bias_dropout_add,
            pre_mlp_layernorm=b.layer_norm(),
            mlp=ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=b.column_parallel_linear(), linear_fc2=b.row_parallel_linear())),
            mlp_bda=get_bias_dropout_add,
        ),
    )


bert_layer_local_spec = get_bert_layer_local_spec()
bert_layer_with_transformer_engine_spec = bert_layer_local_spec
