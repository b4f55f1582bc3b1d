"""Megatron MoE (Mixtral-style) ⇄ Hugging Face ``MixtralForCausalLM`` parameter layout (reference ``tools/checkpoint/loader_mixtral_hf.py`` / saver).

On top of the dense mapping of ``hf_llama.py`` (attention, norms, embeddings) the expert block maps as

    mlp.router.weight [E, h]                          ⇄  block_sparse_moe.gate.weight
    grouped experts  weight1 [E, 2·ffn, h] (gate|up)  ⇄  experts.{e}.w1 (gate) / experts.{e}.w3 (up)
                     weight2 [E, h, ffn]              ⇄  experts.{e}.w2 (down)
    sequential experts  local_experts.{e}.linear_fc1 / linear_fc2  — same split per expert.

Operates on FULL tensors: gather EP/TP shards first (a ``dist_checkpointing`` load with EP = TP = 1 does it)."""
from __future__ import annotations

from typing import Dict

import torch

from .hf_llama import _layer_prefixes, hf_llama_to_megatron, megatron_to_hf_llama


def _dense_part(sd: Dict[str, torch.Tensor], drop: str) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in sd.items() if drop not in k}


def megatron_to_hf_mixtral(sd: Dict[str, torch.Tensor], num_attention_heads: int, num_query_groups: int, kv_channels: int) -> Dict[str, torch.Tensor]:
    layers = _layer_prefixes(sd)
    # reuse the dense converter for everything but the MLP (it expects linear_fc1/2: feed it placeholders and drop them again)
    dense = dict(_dense_part(sd, ".mlp."))
    h = sd["decoder.final_layernorm.weight"].shape[0]
    for i in layers:
        dense[f"decoder.layers.{i}.mlp.linear_fc1.weight"] = torch.empty(2, h)
        dense[f"decoder.layers.{i}.mlp.linear_fc2.weight"] = torch.empty(h, 1)
    out = {k: v for k, v in megatron_to_hf_llama(dense, num_attention_heads, num_query_groups, kv_channels).items() if ".mlp." not in k}
    for i in layers:
        p, q = f"decoder.layers.{i}.mlp.", f"model.layers.{i}.block_sparse_moe."
        out[q + "gate.weight"] = sd[p + "router.weight"]
        if p + "experts.weight1" in sd:
            w1, w2 = sd[p + "experts.weight1"], sd[p + "experts.weight2"]
            per = [(w1[e], w2[e]) for e in range(w1.shape[0])]
        else:
            E = len({k.split(".")[6] for k in sd if k.startswith(p + "experts.local_experts.") and k.endswith("linear_fc1.weight")})
            per = [(sd[f"{p}experts.local_experts.{e}.linear_fc1.weight"], sd[f"{p}experts.local_experts.{e}.linear_fc2.weight"]) for e in range(E)]
        for e, (fc1, fc2) in enumerate(per):
            ffn = fc1.shape[0] // 2
            out[f"{q}experts.{e}.w1.weight"], out[f"{q}experts.{e}.w3.weight"] = fc1[:ffn].clone(), fc1[ffn:].clone()
            out[f"{q}experts.{e}.w2.weight"] = fc2.clone()
    return out


def hf_mixtral_to_megatron(hf: Dict[str, torch.Tensor], num_attention_heads: int, num_query_groups: int, kv_channels: int, grouped: bool = True,
                           tie_embeddings: bool = False) -> Dict[str, torch.Tensor]:
    n = len({k.split(".")[2] for k in hf if k.startswith("model.layers.")})
    h = hf["model.norm.weight"].shape[0]
    dense = {k: v for k, v in hf.items() if "block_sparse_moe" not in k}
    for i in range(n):
        q = f"model.layers.{i}.mlp."
        dense[q + "gate_proj.weight"] = dense[q + "up_proj.weight"] = torch.empty(1, h)
        dense[q + "down_proj.weight"] = torch.empty(h, 1)
    out = {k: v for k, v in hf_llama_to_megatron(dense, num_attention_heads, num_query_groups, kv_channels, tie_embeddings).items() if ".mlp." not in k}
    for i in range(n):
        p, q = f"decoder.layers.{i}.mlp.", f"model.layers.{i}.block_sparse_moe."
        out[p + "router.weight"] = hf[q + "gate.weight"]
        E = len({k.split(".")[5] for k in hf if k.startswith(q + "experts.") and k.endswith("w1.weight")})
        fc1 = [torch.cat([hf[f"{q}experts.{e}.w1.weight"], hf[f"{q}experts.{e}.w3.weight"]], dim=0) for e in range(E)]
        fc2 = [hf[f"{q}experts.{e}.w2.weight"] for e in range(E)]
        if grouped:
            out[p + "experts.weight1"], out[p + "experts.weight2"] = torch.stack(fc1), torch.stack(fc2)
        else:
            for e in range(E):
                out[f"{p}experts.local_experts.{e}.linear_fc1.weight"], out[f"{p}experts.local_experts.{e}.linear_fc2.weight"] = fc1[e], fc2[e]
    return out


def hub_to_fused_experts_layout(hf: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Hub checkpoint names (``block_sparse_moe.experts.{e}.w1/w2/w3``) → the fused in-memory layout of ``transformers`` ≥ 5
    (``mlp.gate.weight``, ``mlp.experts.gate_up_proj [E, 2·ffn, h]``, ``mlp.experts.down_proj [E, h, ffn]``) — what ``load_state_dict`` of a
    freshly constructed ``MixtralForCausalLM`` expects (``from_pretrained`` applies the same conversion itself)."""
    out = {k: v for k, v in hf.items() if "block_sparse_moe" not in k}
    layers = sorted({int(k.split(".")[2]) for k in hf if "block_sparse_moe" in k})
    for i in layers:
        q = f"model.layers.{i}.block_sparse_moe."
        E = len({k.split(".")[5] for k in hf if k.startswith(q + "experts.") and k.endswith("w1.weight")})
        out[f"model.layers.{i}.mlp.gate.weight"] = hf[q + "gate.weight"]
        out[f"model.layers.{i}.mlp.experts.gate_up_proj"] = torch.stack([torch.cat([hf[f"{q}experts.{e}.w1.weight"], hf[f"{q}experts.{e}.w3.weight"]], 0) for e in range(E)])
        out[f"model.layers.{i}.mlp.experts.down_proj"] = torch.stack([hf[f"{q}experts.{e}.w2.weight"] for e in range(E)])
    return out
