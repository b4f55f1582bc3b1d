"""Analytic FLOPs model (same accounting as the reference, ``megatron/training/training.py:802-1390``):
fwd+bwd = 3x fwd, FMA = 2 FLOPs, gated FFN has 3 projections, causal attention counted at 1/2,
LM-head logits included, activation recompute NOT counted."""
from __future__ import annotations


def num_floating_point_operations(*, num_layers, hidden_size, ffn_hidden_size, num_attention_heads, num_query_groups, kv_channels, vocab_size,
                                  seq_length, batch_size, swiglu=True, num_moe_experts=None, moe_router_topk=1, moe_layer_freq=1,
                                  moe_ffn_hidden_size=None, mtp_num_layers=0) -> float:
    s, B, h = seq_length, batch_size, hidden_size
    ffn_hidden_size = ffn_hidden_size or 4 * hidden_size
    kv_channels = kv_channels or hidden_size // num_attention_heads
    num_query_groups = num_query_groups or num_attention_heads
    q_proj = kv_channels * num_attention_heads
    kv_proj = kv_channels * num_query_groups
    gate = 3 if swiglu else 2  # number of [h, ffn]-sized matrices in the FFN
    tokens = B * s
    # per layer, forward, in MACs
    attn_linear = h * (q_proj + 2 * kv_proj) + q_proj * h
    attn_core = 2 * q_proj * s / 2  # QK^T and PV, causal → half
    dense_ffn = gate * h * ffn_hidden_size
    if num_moe_experts:
        mffn = moe_ffn_hidden_size or ffn_hidden_size
        n_moe = num_layers // moe_layer_freq if isinstance(moe_layer_freq, int) else sum(moe_layer_freq)
        n_dense = num_layers - n_moe
        ffn_total = n_dense * dense_ffn + n_moe * (gate * h * mffn * moe_router_topk + h * num_moe_experts)
    else:
        ffn_total = num_layers * dense_ffn
    macs = tokens * (num_layers * (attn_linear + attn_core) + ffn_total + h * vocab_size * (1 + mtp_num_layers))
    return 3 * 2 * macs


def flops_per_token(**kw) -> float:
    kw = dict(kw)
    kw["batch_size"] = 1
    return num_floating_point_operations(**kw) / kw["seq_length"]
