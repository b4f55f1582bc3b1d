"""Convert between this framework's GPT/Llama parameter layout and the Hugging Face ``LlamaForCausalLM`` layout
(reference ``tools/checkpoint/`` loaders/savers + ``core/export/`` weight converters).

Layout differences handled here:
* fused QKV  ``linear_qkv.weight [g·(r+2)·d, h]`` (per KV group: r query heads, 1 key, 1 value)  ⇄  ``q_proj / k_proj / v_proj``
* fused gate/up  ``linear_fc1.weight [2·ffn, h]`` (gate rows first, then up)                       ⇄  ``gate_proj / up_proj``
* RoPE convention: both sides use the half-split ("rotate_half") layout, so no permutation of q/k rows is needed.
Operates on FULL (unsharded) tensors; gather TP shards first (``dist_checkpointing`` load with TP=1 target, or ``core.resharding``)."""
from __future__ import annotations

from typing import Dict

import torch


def _layer_prefixes(sd: Dict[str, torch.Tensor]):
    return sorted({int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers.")})


def megatron_to_hf_llama(sd: Dict[str, torch.Tensor], num_attention_heads: int, num_query_groups: int, kv_channels: int) -> Dict[str, torch.Tensor]:
    g, r, d = num_query_groups, num_attention_heads // num_query_groups, kv_channels
    out = {"model.embed_tokens.weight": sd["embedding.word_embeddings.weight"], "model.norm.weight": sd["decoder.final_layernorm.weight"]}
    out["lm_head.weight"] = sd.get("output_layer.weight", sd["embedding.word_embeddings.weight"])
    for i in _layer_prefixes(sd):
        p, q = f"decoder.layers.{i}.", f"model.layers.{i}."
        qkv = sd[p + "self_attention.linear_qkv.weight"]
        h = qkv.shape[1]
        qkv = qkv.view(g, (r + 2) * d, h)
        out[q + "self_attn.q_proj.weight"] = qkv[:, : r * d].reshape(g * r * d, h).clone()
        out[q + "self_attn.k_proj.weight"] = qkv[:, r * d : (r + 1) * d].reshape(g * d, h).clone()
        out[q + "self_attn.v_proj.weight"] = qkv[:, (r + 1) * d :].reshape(g * d, h).clone()
        out[q + "self_attn.o_proj.weight"] = sd[p + "self_attention.linear_proj.weight"]
        fc1 = sd[p + "mlp.linear_fc1.weight"]
        ffn = fc1.shape[0] // 2
        out[q + "mlp.gate_proj.weight"], out[q + "mlp.up_proj.weight"] = fc1[:ffn].clone(), fc1[ffn:].clone()
        out[q + "mlp.down_proj.weight"] = sd[p + "mlp.linear_fc2.weight"]
        out[q + "input_layernorm.weight"] = sd[p + "input_layernorm.weight"]
        out[q + "post_attention_layernorm.weight"] = sd[p + "pre_mlp_layernorm.weight"]
    return out


def hf_llama_to_megatron(hf: Dict[str, torch.Tensor], num_attention_heads: int, num_query_groups: int, kv_channels: int, tie_embeddings: bool = False) -> Dict[str, torch.Tensor]:
    g, r, d = num_query_groups, num_attention_heads // num_query_groups, kv_channels
    out = {"embedding.word_embeddings.weight": hf["model.embed_tokens.weight"], "decoder.final_layernorm.weight": hf["model.norm.weight"]}
    if not tie_embeddings:
        out["output_layer.weight"] = hf.get("lm_head.weight", hf["model.embed_tokens.weight"])
    n = len({k.split(".")[2] for k in hf if k.startswith("model.layers.")})
    for i in range(n):
        p, q = f"decoder.layers.{i}.", f"model.layers.{i}."
        h = hf[q + "self_attn.q_proj.weight"].shape[1]
        wq = hf[q + "self_attn.q_proj.weight"].view(g, r * d, h)
        wk = hf[q + "self_attn.k_proj.weight"].view(g, d, h)
        wv = hf[q + "self_attn.v_proj.weight"].view(g, d, h)
        out[p + "self_attention.linear_qkv.weight"] = torch.cat([wq, wk, wv], dim=1).reshape(g * (r + 2) * d, h)
        out[p + "self_attention.linear_proj.weight"] = hf[q + "self_attn.o_proj.weight"]
        out[p + "mlp.linear_fc1.weight"] = torch.cat([hf[q + "mlp.gate_proj.weight"], hf[q + "mlp.up_proj.weight"]], dim=0)
        out[p + "mlp.linear_fc2.weight"] = hf[q + "mlp.down_proj.weight"]
        out[p + "input_layernorm.weight"] = hf[q + "input_layernorm.weight"]
        out[p + "pre_mlp_layernorm.weight"] = hf[q + "post_attention_layernorm.weight"]
    return out
