"""Weights + config export in the TensorRT-LLM checkpoint format (reference ``core/export/trtllm/``: ``TRTLLMLayers`` name map,
``single_device`` / ``distributed`` weight converters, ``trt_model_config``).

A TRT-LLM checkpoint is ``config.json`` plus one ``rank{r}.safetensors`` per inference rank with fixed tensor names
(``transformer.layers.{i}.attention.qkv.weight`` …).  Nothing of TRT-LLM is needed to WRITE one — it is a renaming, three layout
changes and a split:

* fused QKV: ours is grouped ``[g · (r + 2) · d, h]`` (per KV group: r query heads, K, V); TRT-LLM wants ``[Q ‖ K ‖ V]`` with all query
  heads first.  When the inference TP is larger than the number of KV groups each K/V head is replicated ``tp / g`` times so every
  rank owns a whole one.
* gated MLP: our ``linear_fc1`` stacks ``[gate ; up]``; TRT-LLM's ``mlp.fc`` is the *activated* (gate) branch and ``mlp.gate`` the
  linear one.
* vocabulary rows are padded to a multiple of 64 · tp.

``TRTLLMWeightsConverter.convert`` works from a FULL (unsharded) state dict — e.g. ``dist_checkpointing`` loaded at TP=1 — and returns
the per-rank dicts; ``DistributedTRTLLMWeightsConverter`` converts the local TP shard in place when training TP == inference TP (no
gather).  ``save_trtllm_checkpoint`` writes the directory (safetensors when the package is importable, ``torch.save`` otherwise)."""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass
from enum import Enum
from typing import Dict, List, Optional, Tuple

import torch


class TRTLLMLayers(str, Enum):
    position_embedding = "transformer.position_embedding.weight"
    vocab_embedding = "transformer.vocab_embedding.weight"
    lm_head = "lm_head.weight"
    final_layernorm_weight = "transformer.ln_f.weight"
    final_layernorm_bias = "transformer.ln_f.bias"
    input_layernorm_weight = "transformer.layers.input_layernorm.weight"
    input_layernorm_bias = "transformer.layers.input_layernorm.bias"
    attention_qkv_weight = "transformer.layers.attention.qkv.weight"
    attention_qkv_bias = "transformer.layers.attention.qkv.bias"
    attention_dense_weight = "transformer.layers.attention.dense.weight"
    attention_dense_bias = "transformer.layers.attention.dense.bias"
    mlp_fc_weight = "transformer.layers.mlp.fc.weight"
    mlp_fc_bias = "transformer.layers.mlp.fc.bias"
    mlp_gate_weight = "transformer.layers.mlp.gate.weight"
    post_layernorm_weight = "transformer.layers.post_layernorm.weight"
    post_layernorm_bias = "transformer.layers.post_layernorm.bias"
    mlp_projection_weight = "transformer.layers.mlp.proj.weight"
    mlp_projection_bias = "transformer.layers.mlp.proj.bias"
    mlp_router_weight = "transformer.layers.mlp.router.weight"
    mlp_fc_weight_mixture_of_experts = "transformer.layers.mlp.fc.weight"
    mlp_projection_weight_mixture_of_experts = "transformer.layers.mlp.proj.weight"

    @staticmethod
    def return_layer_name_and_number(layer_name: str) -> Tuple[str, Optional[int]]:
        """``decoder.layers.2.mlp.linear_fc1.weight`` → (``decoder.layers.mlp.linear_fc1.weight``, 2)."""
        m = re.search(r"(?<=layers\.)\d+(?=\.)", layer_name)
        if not m:
            return layer_name, None
        return layer_name[: m.start()] + layer_name[m.end() + 1:], int(m.group(0))

    def with_layer(self, i: int) -> str:
        return self.value.replace("transformer.layers.", f"transformer.layers.{i}.")


DEFAULT_CONVERSION_DICT: Dict[str, TRTLLMLayers] = {
    "embedding.word_embeddings.weight": TRTLLMLayers.vocab_embedding,
    "embedding.position_embeddings.weight": TRTLLMLayers.position_embedding,
    "output_layer.weight": TRTLLMLayers.lm_head,
    "decoder.final_layernorm.weight": TRTLLMLayers.final_layernorm_weight,
    "decoder.final_layernorm.bias": TRTLLMLayers.final_layernorm_bias,
    "decoder.layers.input_layernorm.weight": TRTLLMLayers.input_layernorm_weight,
    "decoder.layers.input_layernorm.bias": TRTLLMLayers.input_layernorm_bias,
    "decoder.layers.self_attention.linear_qkv.weight": TRTLLMLayers.attention_qkv_weight,
    "decoder.layers.self_attention.linear_qkv.bias": TRTLLMLayers.attention_qkv_bias,
    "decoder.layers.self_attention.linear_proj.weight": TRTLLMLayers.attention_dense_weight,
    "decoder.layers.self_attention.linear_proj.bias": TRTLLMLayers.attention_dense_bias,
    "decoder.layers.pre_mlp_layernorm.weight": TRTLLMLayers.post_layernorm_weight,
    "decoder.layers.pre_mlp_layernorm.bias": TRTLLMLayers.post_layernorm_bias,
    "decoder.layers.mlp.linear_fc1.weight": TRTLLMLayers.mlp_fc_weight,
    "decoder.layers.mlp.linear_fc1.bias": TRTLLMLayers.mlp_fc_bias,
    "decoder.layers.mlp.linear_fc2.weight": TRTLLMLayers.mlp_projection_weight,
    "decoder.layers.mlp.linear_fc2.bias": TRTLLMLayers.mlp_projection_bias,
    "decoder.layers.mlp.router.weight": TRTLLMLayers.mlp_router_weight,
}


@dataclass
class ExportConfig:
    inference_tp_size: int = 1
    inference_pp_size: int = 1
    dtype: torch.dtype = torch.bfloat16
    share_embeddings_and_output_weights: bool = False


def pad_vocab_size(vocab_size: int, tp_size: int) -> int:
    m = 64 * tp_size
    return -(-vocab_size // m) * m


def trtllm_model_config(cfg, vocab_size: int, max_position_embeddings: int, export: ExportConfig, architecture: str = "LlamaForCausalLM",
                        position_embedding_type: str = "rope_gpt_neox", rotary_base: float = 10000.0) -> dict:
    """The ``config.json`` a TRT-LLM ``trtllm-build`` run reads."""
    dt = {torch.bfloat16: "bfloat16", torch.float16: "float16", torch.float32: "float32"}[export.dtype]
    out = {
        "architecture": architecture, "dtype": dt, "logits_dtype": "float32", "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_attention_heads,
        "num_key_value_heads": cfg.num_query_groups or cfg.num_attention_heads, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.ffn_hidden_size,
        "head_size": cfg.kv_channels, "norm_epsilon": cfg.layernorm_epsilon, "vocab_size": pad_vocab_size(vocab_size, export.inference_tp_size),
        "max_position_embeddings": max_position_embeddings, "position_embedding_type": position_embedding_type, "rotary_base": rotary_base,
        "hidden_act": "swiglu" if cfg.gated_linear_unit else getattr(cfg.activation_func, "__name__", "gelu"), "bias": bool(cfg.add_bias_linear),
        "share_embedding_table": export.share_embeddings_and_output_weights, "use_parallel_embedding": export.inference_tp_size > 1, "embedding_sharding_dim": 0,
        "mapping": {"world_size": export.inference_tp_size * export.inference_pp_size, "tp_size": export.inference_tp_size, "pp_size": export.inference_pp_size},
        "quantization": {"quant_algo": None, "kv_cache_quant_algo": None},
    }
    if cfg.num_moe_experts:
        out["moe"] = {"num_experts": cfg.num_moe_experts, "top_k": cfg.moe_router_topk, "normalization_mode": 1}
        out["moe_intermediate_size"] = cfg.moe_ffn_hidden_size or cfg.ffn_hidden_size
    return out


def _regroup_qkv(w: torch.Tensor, g: int, r: int, d: int, tp: int) -> List[torch.Tensor]:
    """Grouped ``[g·(r+2)·d, …]`` → per-rank ``[Q_rank ‖ K_rank ‖ V_rank]``.  Works for weights (2-D) and biases (1-D)."""
    tail = w.shape[1:]
    w = w.reshape(g, (r + 2) * d, *tail)
    q = w[:, : r * d].reshape(g * r, d, *tail)                        # heads in group-major order == head index order
    k, v = w[:, r * d : (r + 1) * d], w[:, (r + 1) * d :]            # [g, d, …]
    if tp > g:
        assert tp % g == 0, f"inference tp ({tp}) must be a multiple of the KV groups ({g})"
        k, v = k.repeat_interleave(tp // g, 0), v.repeat_interleave(tp // g, 0)
    else:
        assert g % tp == 0, f"KV groups ({g}) must be divisible by the inference tp ({tp})"
    out = []
    for rank in range(tp):
        qs, ks, vs = q.chunk(tp, 0)[rank], k.chunk(tp, 0)[rank], v.chunk(tp, 0)[rank]
        out.append(torch.cat([qs.reshape(-1, *tail), ks.reshape(-1, *tail), vs.reshape(-1, *tail)], 0).contiguous())
    return out


def _split(w: torch.Tensor, tp: int, dim: int) -> List[torch.Tensor]:
    return [c.contiguous() for c in w.chunk(tp, dim)]


class TRTLLMWeightsConverter:
    """Full (TP=1) state dict → ``[rank0 weights, rank1 weights, …]`` for ``inference_tp_size`` ranks (single-device converter)."""

    def __init__(self, export: ExportConfig, transformer_config, conversion_dict: Optional[Dict[str, TRTLLMLayers]] = None):
        self.export, self.cfg = export, transformer_config
        self.map = conversion_dict or DEFAULT_CONVERSION_DICT
        self.g = transformer_config.num_query_groups or transformer_config.num_attention_heads
        self.r = transformer_config.num_attention_heads // self.g
        self.d = transformer_config.kv_channels

    def _per_rank(self, name: TRTLLMLayers, val: torch.Tensor, vocab_size: Optional[int]) -> List[torch.Tensor]:
        tp, L = self.export.inference_tp_size, TRTLLMLayers
        val = val.to(self.export.dtype)
        if name in (L.attention_qkv_weight, L.attention_qkv_bias):
            return _regroup_qkv(val, self.g, self.r, self.d, tp)
        if name in (L.vocab_embedding, L.lm_head):
            rows = pad_vocab_size(vocab_size or val.shape[0], tp)
            if val.shape[0] < rows:
                val = torch.cat([val, val.new_zeros(rows - val.shape[0], val.shape[1])])
            return _split(val[:rows], tp, 0)
        if name in (L.mlp_fc_weight, L.mlp_fc_bias, L.mlp_gate_weight):
            return _split(val, tp, 0)
        if name in (L.attention_dense_weight, L.mlp_projection_weight):
            return _split(val, tp, 1)
        return [val] * tp                                             # norms, row-parallel biases, router, position embedding

    def convert(self, state_dict: Dict[str, torch.Tensor], vocab_size: Optional[int] = None) -> List[Dict[str, torch.Tensor]]:
        tp, L = self.export.inference_tp_size, TRTLLMLayers
        ranks: List[Dict[str, torch.Tensor]] = [dict() for _ in range(tp)]

        def put(name: TRTLLMLayers, layer: Optional[int], val: torch.Tensor, key_override: Optional[str] = None):
            key = key_override or (name.with_layer(layer) if layer is not None else name.value)
            for rk, t in enumerate(self._per_rank(name, val, vocab_size)):
                ranks[rk][key] = t

        experts: Dict[Tuple[int, str], Dict[int, torch.Tensor]] = {}
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or k.endswith("_extra_state"):
                continue
            m = re.match(r"decoder\.layers\.(\d+)\.mlp\.experts\.(?:local_experts\.(\d+)\.linear_fc([12])\.weight|weight([12]))$", k)
            if m:
                layer = int(m.group(1))
                if m.group(2) is not None:
                    experts.setdefault((layer, m.group(3)), {})[int(m.group(2))] = v
                else:
                    for e in range(v.shape[0]):
                        experts.setdefault((layer, m.group(4)), {})[e] = v[e]
                continue
            base, layer = L.return_layer_name_and_number(k)
            name = self.map.get(base)
            if name is None:
                continue
            if name is L.mlp_fc_weight and self.cfg.gated_linear_unit:
                gate, up = v.chunk(2, 0)
                put(L.mlp_fc_weight, layer, gate)
                put(L.mlp_gate_weight, layer, up)
            else:
                put(name, layer, v)
        # experts: stacked [E, …] tensors, each expert's FFN dimension split over TP
        for (layer, which), per in experts.items():
            ws = [per[e].to(self.export.dtype) for e in sorted(per)]
            if which == "1":
                if self.cfg.gated_linear_unit:                         # TRT-LLM MoE fc = [up ; gate] per expert
                    per_rank = [torch.stack([torch.cat([w.chunk(2, 0)[1].chunk(tp, 0)[rk], w.chunk(2, 0)[0].chunk(tp, 0)[rk]], 0) for w in ws]) for rk in range(tp)]
                else:
                    per_rank = [torch.stack([w.chunk(tp, 0)[rk] for w in ws]) for rk in range(tp)]
                key = L.mlp_fc_weight_mixture_of_experts.with_layer(layer)
            else:
                per_rank = [torch.stack([w.chunk(tp, 1)[rk] for w in ws]) for rk in range(tp)]
                key = L.mlp_projection_weight_mixture_of_experts.with_layer(layer)
            for rk in range(tp):
                ranks[rk][key] = per_rank[rk].contiguous()
        if self.export.share_embeddings_and_output_weights or L.lm_head.value not in ranks[0]:
            for rk in range(tp):
                ranks[rk][L.lm_head.value] = ranks[rk][L.vocab_embedding.value]
        return ranks


class DistributedTRTLLMWeightsConverter(TRTLLMWeightsConverter):
    """Training TP == inference TP: every rank renames / re-lays-out ITS shard, nothing is gathered.  Our TP shard of the fused QKV holds
    whole KV groups (``g / tp`` of them) and of ``linear_fc1`` holds matching ``[gate ; up]`` halves, so the single-rank routine applies."""

    def __init__(self, export: ExportConfig, transformer_config, tp_rank: int, tp_size: int, conversion_dict=None):
        super().__init__(ExportConfig(1, export.inference_pp_size, export.dtype, export.share_embeddings_and_output_weights), transformer_config, conversion_dict)
        assert export.inference_tp_size == tp_size, "distributed conversion keeps the training TP"
        self.tp_rank, self.tp_size = tp_rank, tp_size
        assert self.g % tp_size == 0
        self.g //= tp_size

    def convert(self, state_dict, vocab_size: Optional[int] = None):
        local_vocab = None if vocab_size is None else pad_vocab_size(vocab_size, self.tp_size) // self.tp_size
        return super().convert(state_dict, local_vocab)[0]


def save_trtllm_checkpoint(out_dir: str, rank_weights: List[Dict[str, torch.Tensor]], config: dict) -> None:
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "config.json"), "w") as f:
        json.dump(config, f, indent=2)
    try:
        from safetensors.torch import save_file

        for r, w in enumerate(rank_weights):
            save_file({k: v.contiguous().clone() for k, v in w.items()}, os.path.join(out_dir, f"rank{r}.safetensors"))
    except ImportError:
        for r, w in enumerate(rank_weights):
            torch.save(w, os.path.join(out_dir, f"rank{r}.pt"))
