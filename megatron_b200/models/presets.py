"""Named model configurations (the five BASELINE.json configs + small test models)."""
from __future__ import annotations

from typing import Dict

import torch.nn.functional as F

PRESETS: Dict[str, dict] = {
    # Llama-3 8B: examples/llama/train_llama3_8b_h100_fp8.sh MODEL_ARGS in the reference
    "llama3_8b": dict(num_layers=32, hidden_size=4096, ffn_hidden_size=14336, num_attention_heads=32, num_query_groups=8, kv_channels=128,
                      vocab_size=128256, seq_length=8192, normalization="RMSNorm", swiglu=True, rotary_base=500000, untie=True, bias=False),
    "llama3_8b_dp": dict(num_layers=32, hidden_size=4096, ffn_hidden_size=14336, num_attention_heads=32, num_query_groups=8, kv_channels=128,
                         vocab_size=128256, seq_length=8192, normalization="RMSNorm", swiglu=True, rotary_base=500000, untie=True, bias=False),   # same model; a second name so that bench.py routes it through the generic (DP / PP / EP) arm
    "llama3_70b": dict(num_layers=80, hidden_size=8192, ffn_hidden_size=28672, num_attention_heads=64, num_query_groups=8, kv_channels=128,
                       vocab_size=128256, seq_length=8192, normalization="RMSNorm", swiglu=True, rotary_base=500000, untie=True, bias=False),
    "gpt3_6.7b": dict(num_layers=32, hidden_size=4096, ffn_hidden_size=16384, num_attention_heads=32, num_query_groups=32, kv_channels=128,
                      vocab_size=50304, seq_length=2048, normalization="LayerNorm", swiglu=False, rotary_base=None, untie=False, bias=True),
    "mixtral_8x7b": dict(num_layers=32, hidden_size=4096, ffn_hidden_size=14336, num_attention_heads=32, num_query_groups=8, kv_channels=128,
                         vocab_size=32000, seq_length=4096, normalization="RMSNorm", swiglu=True, rotary_base=1000000, untie=True, bias=False,
                         num_moe_experts=8, moe_router_topk=2),
    "gpt2_small": dict(num_layers=12, hidden_size=768, ffn_hidden_size=3072, num_attention_heads=12, num_query_groups=12, kv_channels=64,
                       vocab_size=50304, seq_length=1024, normalization="LayerNorm", swiglu=False, rotary_base=None, untie=False, bias=True),
    "tiny_mixtral": dict(num_layers=2, hidden_size=256, ffn_hidden_size=512, num_attention_heads=4, num_query_groups=2, kv_channels=64,
                         vocab_size=1024, seq_length=256, normalization="RMSNorm", swiglu=True, rotary_base=10000, untie=True, bias=False,
                         num_moe_experts=4, moe_router_topk=2),
    "tiny_llama": dict(num_layers=2, hidden_size=256, ffn_hidden_size=512, num_attention_heads=4, num_query_groups=2, kv_channels=64,
                       vocab_size=1024, seq_length=256, normalization="RMSNorm", swiglu=True, rotary_base=10000, untie=True, bias=False),
}


def make_transformer_config(name: str, **overrides):
    """TransformerConfig for a preset; ``overrides`` may set parallel sizes, dtype, num_layers …"""
    from ..core.transformer.transformer_config import TransformerConfig

    p = dict(PRESETS[name])
    for k in list(overrides):
        if k in p:
            p[k] = overrides.pop(k)
    kw = dict(
        num_layers=p["num_layers"], hidden_size=p["hidden_size"], ffn_hidden_size=p["ffn_hidden_size"],
        num_attention_heads=p["num_attention_heads"], num_query_groups=p["num_query_groups"], kv_channels=p["kv_channels"],
        normalization=p["normalization"], gated_linear_unit=p["swiglu"], activation_func=F.silu if p["swiglu"] else F.gelu,
        add_bias_linear=p["bias"], hidden_dropout=0.0, attention_dropout=0.0, bias_activation_fusion=True, bias_dropout_fusion=True,
        apply_rope_fusion=True, masked_softmax_fusion=True,
    )
    if p.get("num_moe_experts"):
        kw.update(num_moe_experts=p["num_moe_experts"], moe_router_topk=p["moe_router_topk"], moe_token_dispatcher_type="alltoall",
                  moe_router_load_balancing_type="aux_loss", moe_aux_loss_coeff=1e-2, moe_grouped_gemm=True)
    kw.update(overrides)
    return TransformerConfig(**kw), p


def build_gpt_model(name: str, pre_process=True, post_process=True, vp_stage=None, **overrides):
    from ..core.models.gpt.gpt_layer_specs import get_gpt_decoder_block_spec, get_gpt_layer_local_spec
    from ..core.models.gpt.gpt_model import GPTModel

    cfg, p = make_transformer_config(name, **overrides)
    if cfg.num_moe_experts is not None:
        spec = get_gpt_decoder_block_spec(cfg, vp_stage=vp_stage)
    else:
        spec = get_gpt_layer_local_spec(normalization=cfg.normalization, qk_layernorm=cfg.qk_layernorm)
    model = GPTModel(
        cfg, spec, vocab_size=p["vocab_size"], max_sequence_length=p["seq_length"], pre_process=pre_process, post_process=post_process,
        parallel_output=True, share_embeddings_and_output_weights=not p["untie"],
        position_embedding_type="rope" if p["rotary_base"] else "learned_absolute", rotary_base=p["rotary_base"] or 10000, vp_stage=vp_stage,
    )
    return model, cfg, p
