"""Analytic FLOPs model (same accounting as the reference, ``megatron/training/training.py:802-1390``):
fwd+bwd = 3x fwd, FMA = 2 FLOPs, gated FFN has 3 projections, causal attention counted at 1/2,
LM-head logits included, activation recompute NOT counted.

Structure (ours): every layer type contributes ``(token-linear MACs per token, core-attention MACs per (query, key) PAIR)``; a batch contributes
``tokens`` (real, unpadded) and ``pairs`` (causal pairs actually attended: ``sum L_i^2 / 2`` for packed / THD batches, the sliding-window wedge when a
window is set).  Covered: MHA / GQA (+ output gate), multi-latent attention (q / kv LoRA ranks, decoupled RoPE dims), gated-delta-net linear attention
layers interleaved by ``linear_attention_freq``, dense and MoE FFNs (routed top-k, shared expert, latent experts, router), multi-token-prediction layers,
hybrid Mamba-2 stacks, sliding-window attention, THD batches."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Union


# ---- batch geometry -------------------------------------------------------------------------------------------------------------------
def causal_pairs(seq_len: float, window: Optional[int] = None) -> float:
    """(query, key) pairs of one causal sequence, counted the reference's way (L^2 / 2); with a sliding window of W keys each query sees at most W keys:
    the wedge W*L - W^2/2 for L > W."""
    if window is None or window <= 0 or seq_len <= window:
        return seq_len * seq_len / 2.0
    return window * seq_len - window * window / 2.0


def batch_geometry(seq_length: int, batch_size: int, seq_lens: Optional[Iterable[int]] = None, window: Optional[int] = None,
                   seqlen_squared_sum_in_batch: Optional[float] = None, total_real_tokens_in_batch: Optional[float] = None):
    """→ (real tokens, causal pairs).  ``seq_lens``: the REAL sub-sequence lengths of a packed (THD) global batch; or pass the two sums the reference's
    signature uses (``sum L_i^2`` and ``sum L_i``)."""
    if seq_lens is not None:
        lens = list(seq_lens)
        return float(sum(lens)), float(sum(causal_pairs(l, window) for l in lens))
    tokens = float(total_real_tokens_in_batch) if total_real_tokens_in_batch is not None else float(batch_size) * seq_length
    if seqlen_squared_sum_in_batch is not None:
        pairs = float(seqlen_squared_sum_in_batch) / 2.0
        if window:
            pairs = min(pairs, tokens * window)          # upper bound without the per-sequence lengths
        return tokens, pairs
    return tokens, float(batch_size) * causal_pairs(seq_length, window)


# ---- per-layer terms (MACs) -----------------------------------------------------------------------------------------------------------
def attention_macs(hidden_size: int, num_attention_heads: int, num_query_groups: Optional[int], kv_channels: Optional[int], attention_output_gate: bool = False):
    """→ (projection MACs per token, core MACs per pair) of one MHA / GQA layer."""
    d = kv_channels or hidden_size // num_attention_heads
    g = num_query_groups or num_attention_heads
    q, kv = d * num_attention_heads, d * g
    proj = hidden_size * (q + 2 * kv + (q if attention_output_gate else 0)) + q * hidden_size
    core = 2 * q                                        # QK^T and PV: d MACs per head each
    return proj, core


def mla_macs(hidden_size: int, num_attention_heads: int, q_lora_rank: Optional[int], kv_lora_rank: int, qk_head_dim: int, qk_pos_emb_head_dim: int, v_head_dim: int):
    """Multi-latent attention (reference :1086-1136): optional low-rank Q, compressed KV latent + shared RoPE key, per-head up-projections."""
    qk = qk_head_dim + qk_pos_emb_head_dim
    if q_lora_rank is None:
        q_term = hidden_size * num_attention_heads * qk
    else:
        q_term = q_lora_rank * (hidden_size + num_attention_heads * qk + 1)
    kv_term = kv_lora_rank * (hidden_size + num_attention_heads * (qk_head_dim + v_head_dim) + 1) + hidden_size * qk_pos_emb_head_dim
    o_term = num_attention_heads * v_head_dim * hidden_size
    core = num_attention_heads * (qk + v_head_dim)
    return q_term + kv_term + o_term, core


def gated_delta_net_macs(hidden_size: int, qk_head_dim: int = 128, v_head_dim: int = 128, num_qk_heads: int = 16, num_v_heads: int = 32, conv_kernel_dim: int = 4,
                         gdn2: bool = False) -> float:
    """Linear-attention layer (no L^2 term): in-proj, short conv, the delta rule's four d_v x d_v products per head, out-proj."""
    qk, v = qk_head_dim * num_qk_heads, v_head_dim * num_v_heads
    in_proj = (4 * qk + 3 * v) if gdn2 else (2 * qk + 2 * v + 2 * num_v_heads)
    return hidden_size * in_proj + conv_kernel_dim * (2 * qk + v) + num_v_heads * v_head_dim ** 2 * 4 + hidden_size * v


def mamba_macs(hidden_size: int, state_dim: int = 128, head_dim: int = 64, num_groups: int = 8, num_heads: Optional[int] = None) -> float:
    """Mamba-2 mixer: in-proj (z, x, B, C, dt), the scan (7 flops per (token, channel, state) → 3.5 MACs), out-proj."""
    d_in = 2 * hidden_size
    nheads = num_heads or d_in // head_dim
    return hidden_size * (2 * d_in + 2 * num_groups * state_dim + nheads) + 3.5 * d_in * state_dim + d_in * hidden_size


def dense_ffn_macs(hidden_size: int, ffn_hidden_size: int, swiglu: bool) -> float:
    return (3 if swiglu else 2) * hidden_size * ffn_hidden_size


def moe_ffn_macs(hidden_size: int, moe_ffn_hidden_size: int, topk: int, num_experts: int, swiglu: bool, shared_expert_ffn_hidden_size: int = 0,
                 moe_latent_size: Optional[int] = None) -> float:
    gate = 3 if swiglu else 2
    if moe_latent_size is None:
        routed = gate * hidden_size * moe_ffn_hidden_size * topk
    else:                                               # experts run in a latent space: down/up projections around them
        routed = gate * moe_latent_size * moe_ffn_hidden_size * topk + 2 * hidden_size * moe_latent_size
    return routed + gate * hidden_size * shared_expert_ffn_hidden_size      # (the reference does not count the router GEMM either)


def _pattern(freq: Union[int, Sequence[int], None], n: int, every_kth_is_zero: bool = False) -> List[int]:
    if freq is None:
        return [0] * n
    if isinstance(freq, int):
        return [0 if ((i + 1) % freq == 0) else 1 for i in range(n)] if every_kth_is_zero else [1 if i % freq == 0 else 0 for i in range(n)]
    assert len(freq) == n, f"pattern of length {len(freq)} for {n} layers"
    return list(freq)


def num_floating_point_operations(*, num_layers, hidden_size, ffn_hidden_size, num_attention_heads, num_query_groups, kv_channels, vocab_size,
                                  seq_length, batch_size, swiglu=True, num_moe_experts=None, moe_router_topk=1, moe_layer_freq=1,
                                  moe_ffn_hidden_size=None, mtp_num_layers=0, moe_shared_expert_intermediate_size=None, moe_latent_size=None,
                                  multi_latent_attention=False, q_lora_rank=None, kv_lora_rank=512, qk_head_dim=128, qk_pos_emb_head_dim=64, v_head_dim=128,
                                  attention_output_gate=False, window_size=None, seq_lens=None, seqlen_squared_sum_in_batch=None, total_real_tokens_in_batch=None,
                                  linear_attention_freq=None, linear_attention_kwargs=None, hybrid_layer_counts=None, mamba_kwargs=None) -> float:
    """FLOPs of one global batch (forward + backward).

    ``hybrid_layer_counts``: ``{"attention": a, "mamba": m, "mlp": d, "moe": e}`` for hybrid (Mamba) stacks — then ``num_layers`` / ``moe_layer_freq`` are
    ignored.  ``linear_attention_freq``: int k (every k-th layer is standard attention, the others gated-delta-net) or a 0/1 list (1 = linear)."""
    h = hidden_size
    ffn_hidden_size = ffn_hidden_size or 4 * h
    window = window_size[0] if isinstance(window_size, (tuple, list)) else window_size
    tokens, pairs = batch_geometry(seq_length, batch_size, seq_lens, window, seqlen_squared_sum_in_batch, total_real_tokens_in_batch)
    if multi_latent_attention:
        attn_proj, attn_core = mla_macs(h, num_attention_heads, q_lora_rank, kv_lora_rank, qk_head_dim, qk_pos_emb_head_dim, v_head_dim)
    else:
        attn_proj, attn_core = attention_macs(h, num_attention_heads, num_query_groups, kv_channels, attention_output_gate)
    dense = dense_ffn_macs(h, ffn_hidden_size, swiglu)
    moe = moe_ffn_macs(h, moe_ffn_hidden_size or ffn_hidden_size, moe_router_topk, num_moe_experts or 0, swiglu, moe_shared_expert_intermediate_size or 0,
                       moe_latent_size) if num_moe_experts else 0.0
    mtp = mtp_num_layers or 0
    if hybrid_layer_counts is not None:
        c = hybrid_layer_counts
        linear = c.get("attention", 0) * attn_proj + c.get("mamba", 0) * mamba_macs(h, **(mamba_kwargs or {})) + c.get("mlp", 0) * dense + c.get("moe", 0) * moe
        core = c.get("attention", 0) * attn_core
    else:
        n_total = num_layers + mtp                    # MTP layers repeat the last layer's type
        moe_pat = _pattern(moe_layer_freq, num_layers) if num_moe_experts else [0] * num_layers
        n_moe = sum(moe_pat) + (mtp if (num_moe_experts and moe_pat[-1]) else 0)
        n_dense = n_total - n_moe
        lin_pat = _pattern(linear_attention_freq, n_total, every_kth_is_zero=True)
        n_linear = sum(lin_pat)
        n_std = n_total - n_linear
        linear = n_std * attn_proj + n_linear * gated_delta_net_macs(h, **(linear_attention_kwargs or {})) + n_dense * dense + n_moe * moe
        core = n_std * attn_core
        if mtp:                                        # per MTP depth: norms + the [2h -> h] projection
            linear += mtp * (2 * h * h + 3 * h)          # eh-projection [2h -> h] + three norms (reference :1288-1294)
    linear += h * vocab_size * (1 + mtp)               # logits (and one more per MTP depth)
    return 3 * 2 * (tokens * linear + pairs * core)


def flops_per_token(**kw) -> float:
    kw = dict(kw)
    kw["batch_size"] = 1
    return num_floating_point_operations(**kw) / kw["seq_length"]
