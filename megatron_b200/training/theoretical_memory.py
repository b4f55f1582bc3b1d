"""Analytic per-GPU memory model (reference ``megatron/training/theoretical_memory_usage.py``), calibrated
against measured numbers on B200 (Llama-3 8B TP=1: predicted vs measured 147 GiB peak)."""
from __future__ import annotations

from ..models.presets import PRESETS

GiB = 2**30


def param_counts(p: dict):
    h, f, L, V = p["hidden_size"], p["ffn_hidden_size"], p["num_layers"], p["vocab_size"]
    q = p["kv_channels"] * p["num_attention_heads"]
    kv = p["kv_channels"] * p["num_query_groups"]
    attn = h * (q + 2 * kv) + q * h
    gate = 3 if p["swiglu"] else 2
    E = p.get("num_moe_experts") or 0
    dense_ffn = gate * h * f
    expert = E * dense_ffn + h * E if E else 0
    per_layer_dense = attn + (0 if E else dense_ffn) + 2 * h
    emb = V * h * (2 if p["untie"] else 1)
    return per_layer_dense * L, expert * L, emb


def report(model: str, tp=1, pp=1, dp=1, ep=1, micro_batch=1, seq=None, fp32_grads=False, dist_opt=True, recompute="selective") -> str:
    p = PRESETS[model]
    s = seq or p["seq_length"]
    dense, expert, emb = param_counts(p)
    n_dense = dense / (tp * pp) + emb / tp / (2 if pp > 1 and p["untie"] else 1)
    n_expert = expert / (ep * pp)
    n = n_dense + n_expert
    gb = 4 if fp32_grads else 2
    shard = dp if dist_opt else 1
    static = n * (2 + gb) + (n_dense / shard + n_expert / max(shard // ep, 1)) * 12
    h, f = p["hidden_size"], p["ffn_hidden_size"]
    tok = s * micro_batch
    per_layer = {"none": tok * (h * 2 * 11 + f * 2 * 3) / tp, "selective": tok * (h * 2 * 8.5 + f * 2 * 2) / tp, "full": tok * h * 2 / tp}[recompute]
    act = per_layer * (p["num_layers"] / pp) + tok * p["vocab_size"] / tp * 2
    lines = [
        f"model {model}: {(dense + expert + emb) / 1e9:.2f} B parameters; layout tp={tp} pp={pp} dp={dp} ep={ep}",
        f"  parameters on this GPU        : {n / 1e9:8.3f} B",
        f"  weights + grads               : {n * (2 + gb) / GiB:8.2f} GiB (bf16 weights, {'fp32' if fp32_grads else 'bf16'} main grads)",
        f"  optimizer (fp32 master, m, v) : {(static - n * (2 + gb)) / GiB:8.2f} GiB ({'sharded over dp' if dist_opt else 'replicated'})",
        f"  activations ({recompute:9s})   : {act / GiB:8.2f} GiB (micro-batch {micro_batch} x seq {s})",
        f"  total                         : {(static + act) / GiB:8.2f} GiB of 178 GiB usable on a B200",
    ]
    return "\n".join(lines)
