"""Self/cross attention with fused QKV projection and GQA
(reference ``transformer/attention.py:1279-1954``).

Data flow per TP rank:  x[s/tp,b,h] --AG+GEMM--> qkv[s,b,(g·(q_per_g+2))·d]
--split/RoPE--> flash attention (sm_100a) --GEMM+RS--> out[s/tp,b,h].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch

from ... import ops
from ..enums import AttnMaskType
from ..utils import divide, get_pg_size, get_tensor_model_parallel_group_if_none
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import TransformerConfig


@dataclass
class SelfAttentionSubmodules:
    linear_qkv: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None
    q_layernorm: Union[ModuleSpec, type] = None
    k_layernorm: Union[ModuleSpec, type] = None


@dataclass
class CrossAttentionSubmodules:
    linear_q: Union[ModuleSpec, type] = None
    linear_kv: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None


def packed_positions(cu_seqlens: torch.Tensor, total: int) -> torch.Tensor:
    """Position of every token inside its own sequence, for a pack described by cumulative lengths (tokens after the last boundary continue the last sequence)."""
    cu = cu_seqlens.to(dtype=torch.long)
    idx = torch.arange(total, device=cu.device)
    sid = torch.bucketize(idx, cu[1:], right=True).clamp_(max=cu.numel() - 2)
    return idx - cu[sid]


def _packed_angles(angles: torch.Tensor, psp, total: int, which: str) -> torch.Tensor:
    cu = getattr(psp, f"cu_seqlens_{which}_padded", None)
    if cu is None:
        cu = getattr(psp, f"cu_seqlens_{which}")
    cache = psp.__dict__.setdefault("_angle_rows", {})
    key = (which, total, angles.data_ptr())
    if key not in cache:
        pos = packed_positions(cu.to(angles.device), total)
        assert int(pos.max()) < angles.shape[0], f"rotary table has {angles.shape[0]} rows, the longest packed sequence needs {int(pos.max()) + 1}"
        if len(cache) >= 4:                         # q and kv entries of the current rotary table; anything older belongs to a previous table
            cache.clear()
        cache[key] = angles[pos]
    return cache[key]


class Attention(MegatronModule):
    """Shared machinery: KV cache for inference, RoPE, core attention, output projection."""

    def __init__(self, config: TransformerConfig, submodules, layer_number: int, attn_mask_type: AttnMaskType,
                 attention_type: str, cp_comm_type: Optional[str] = None, pg_collection=None):
        super().__init__(config)
        self.layer_number = layer_number
        self.attn_mask_type = attn_mask_type
        self.attention_type = attention_type
        self.pg_collection = pg_collection
        self.tp_group = pg_collection.tp if pg_collection is not None and getattr(pg_collection, "tp", None) is not None else get_tensor_model_parallel_group_if_none(None)
        ws = get_pg_size(self.tp_group)
        self.query_projection_size = config.kv_channels * config.num_attention_heads
        self.kv_projection_size = config.kv_channels * config.num_query_groups
        self.hidden_size_per_attention_head = divide(self.query_projection_size, config.num_attention_heads)
        self.num_attention_heads_per_partition = divide(config.num_attention_heads, ws)
        if config.num_query_groups < ws:
            # replicate KV heads so every TP rank owns one (reference attention.py GQA<tp path)
            self.num_query_groups_per_partition = 1
        else:
            self.num_query_groups_per_partition = divide(config.num_query_groups, ws)
        self.core_attention = build_module(
            submodules.core_attention, config=config, layer_number=layer_number, attn_mask_type=attn_mask_type,
            attention_type=attention_type, cp_comm_type=cp_comm_type, softmax_scale=config.softmax_scale, pg_collection=pg_collection,
        )
        self.checkpoint_core_attention = config.recompute_granularity == "selective" and "core_attn" in (config.recompute_modules or [])
        self.linear_proj = build_module(
            submodules.linear_proj, self.query_projection_size, config.hidden_size, config=config,
            init_method=config.output_layer_init_method, bias=config.add_bias_linear, input_is_parallel=True,
            skip_bias_add=True, is_expert=False, tp_comm_buffer_name="proj", tp_group=self.tp_group,
        )
        if config.attention_output_gate:
            raise NotImplementedError("attention_output_gate")

    # ---- inference KV cache ----------------------------------------------------
    def _allocate_memory(self, length, batch, dtype, device):
        return torch.empty(length, batch, self.num_query_groups_per_partition, self.hidden_size_per_attention_head, dtype=dtype, device=device)

    def _adjust_key_value_for_inference(self, inference_context, query, key, value, rotary_pos_emb):
        """Append this step's K/V into the static cache and return the valid prefix."""
        if inference_context is None:
            return query, key, value, rotary_pos_emb, self.attn_mask_type
        kvd = inference_context.key_value_memory_dict
        if self.layer_number not in kvd:
            kvd[self.layer_number] = (
                self._allocate_memory(inference_context.max_sequence_length, inference_context.max_batch_size, key.dtype, key.device),
                self._allocate_memory(inference_context.max_sequence_length, inference_context.max_batch_size, value.dtype, value.device),
            )
        kc, vc = kvd[self.layer_number]
        b0 = inference_context.batch_size_offset
        b1 = b0 + key.size(1)
        s0 = inference_context.sequence_len_offset
        s1 = s0 + key.size(0)
        assert s1 <= kc.size(0) and b1 <= kc.size(1), "KV cache overflow"
        kc[s0:s1, b0:b1] = key
        vc[s0:s1, b0:b1] = value
        key, value = kc[:s1, b0:b1], vc[:s1, b0:b1]
        mask_type = self.attn_mask_type
        if rotary_pos_emb is not None:
            q_pos, k_pos = rotary_pos_emb
            rotary_pos_emb = (q_pos[s0:s1], k_pos[:s1])
        if s0 > 0 and query.size(0) == 1:
            mask_type = AttnMaskType.no_mask
        return query, key, value, rotary_pos_emb, mask_type

    def _batched_paged_decode(self, ctx, query, key, value, rotary_pos_emb):
        """One new token for each of B requests with DIFFERENT histories (continuous batching): ``query/key/value [1, B, heads, d]``.
        Every request is rotated at its own position, its K/V entry is written to its page, and it attends to its own prefix —
        one masked attention over the block-table gather instead of B separate forwards (the GEMMs around it see B rows)."""
        assert query.size(0) == 1, "batched decode advances every request by exactly one token"
        pos = ctx.lengths
        if rotary_pos_emb is not None:
            q_pos, k_pos = rotary_pos_emb
            # requests along the "sequence" axis so each row gets its own angle
            query = ops.apply_rope(query.transpose(0, 1).contiguous(), q_pos[pos], self.config.rotary_interleaved).transpose(0, 1)
            key = ops.apply_rope(key.transpose(0, 1).contiguous(), k_pos[pos], self.config.rotary_interleaved).transpose(0, 1)
        li = ctx.layer_index[self.layer_number]
        scale = getattr(self.core_attention, "softmax_scale", None) or query.size(3) ** -0.5
        # fused KV append (one kernel for K and V through the block table) + flash-decoding over the pages: no gather of the history, no [B, L] mask
        table = getattr(ctx, "block_table_i32", ctx.block_table)
        ops.paged_kv_append(key[0], value[0], ctx.cache.k[li], ctx.cache.v[li], table, getattr(ctx, "positions_i32", pos))
        out = ops.paged_attention_decode(query[0], ctx.cache.k[li], ctx.cache.v[li], table, getattr(ctx, "lengths_incl_i32", pos + 1), scale, ctx.max_len)
        return out.reshape(out.shape[0], -1).unsqueeze(0)

    def get_query_key_value_tensors(self, hidden_states, key_value_states=None):
        raise NotImplementedError

    def forward(self, hidden_states, attention_mask, key_value_states=None, inference_context=None, rotary_pos_emb=None,
                rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None, packed_seq_params=None, sequence_len_offset=None,
                *, inference_params=None):
        inference_context = inference_context or inference_params
        query, key, value = self.get_query_key_value_tensors(hidden_states, key_value_states)
        if rotary_pos_emb is not None and not isinstance(rotary_pos_emb, tuple):
            rotary_pos_emb = (rotary_pos_emb, rotary_pos_emb)
        if inference_context is not None and getattr(inference_context, "is_batched_decode", False):
            core_out = self._batched_paged_decode(inference_context, query, key, value, rotary_pos_emb)
            return self.linear_proj(core_out)
        if inference_context is not None and getattr(inference_context, "is_paged_prefill", False):
            # prompt with no cached prefix: ordinary causal attention; the rotated K / V go straight into the request's pages
            if rotary_pos_emb is not None:
                q_pos, k_pos = rotary_pos_emb
                n = key.size(0)
                query = ops.apply_rope(query, q_pos[:n], self.config.rotary_interleaved)
                key = ops.apply_rope(key, k_pos[:n], self.config.rotary_interleaved)
            inference_context.store(self.layer_number, key, value)
            core_out = self.core_attention(query, key, value, attention_mask, attn_mask_type=self.attn_mask_type, attention_bias=attention_bias, packed_seq_params=packed_seq_params)
            return self.linear_proj(core_out)
        n_new = key.size(0)
        query, key_c, value_c, rotary_pos_emb, mask_type = self._adjust_key_value_for_inference(inference_context, query, key, value, rotary_pos_emb)
        if rotary_pos_emb is not None:
            q_pos, k_pos = rotary_pos_emb
            if inference_context is None and packed_seq_params is not None and getattr(packed_seq_params, "qkv_format", "thd") == "thd":
                # packed (THD) batch: positions restart at every sequence boundary — gather the angle rows by per-token position, ONE rope call per tensor
                # (reference: TE's thd rope kernels driven by cu_seqlens, rope_utils.py:180-260)
                q_pos, k_pos = _packed_angles(q_pos, packed_seq_params, query.size(0), "q"), _packed_angles(k_pos, packed_seq_params, key_c.size(0), "kv")
                query = ops.apply_rope(query, q_pos, self.config.rotary_interleaved)
                key_c = ops.apply_rope(key_c, k_pos, self.config.rotary_interleaved)
            elif inference_context is None:
                query = ops.apply_rope(query, q_pos, self.config.rotary_interleaved)
                key_c = ops.apply_rope(key_c, k_pos, self.config.rotary_interleaved)
            else:
                # rotate only the new positions, then refresh them in the cache
                s0 = inference_context.sequence_len_offset
                query = ops.apply_rope(query, q_pos, self.config.rotary_interleaved)
                newk = ops.apply_rope(key, k_pos[s0 : s0 + n_new], self.config.rotary_interleaved)
                kc, _ = inference_context.key_value_memory_dict[self.layer_number]
                b0 = inference_context.batch_size_offset
                kc[s0 : s0 + n_new, b0 : b0 + key.size(1)] = newk
                key_c = kc[: s0 + n_new, b0 : b0 + key.size(1)]
        if self.checkpoint_core_attention and self.training:
            from ..tensor_parallel.random import checkpoint

            def run(q, k, v):
                return self.core_attention(q, k, v, attention_mask, attn_mask_type=mask_type, attention_bias=attention_bias, packed_seq_params=packed_seq_params)

            core_out = checkpoint(run, False, query, key_c, value_c)
        else:
            core_out = self.core_attention(query, key_c, value_c, attention_mask, attn_mask_type=mask_type, attention_bias=attention_bias, packed_seq_params=packed_seq_params)
        output, bias = self.linear_proj(core_out)
        return output, bias


class SelfAttention(Attention):
    """QKV from one fused column-parallel GEMM; per-group layout ``[q_0..q_{r-1}, k, v]``."""

    def __init__(self, config: TransformerConfig, submodules: SelfAttentionSubmodules, layer_number: int,
                 attn_mask_type=AttnMaskType.padding, cp_comm_type: Optional[str] = None, pg_collection=None):
        super().__init__(config, submodules, layer_number, attn_mask_type, "self", cp_comm_type, pg_collection)
        self.linear_qkv_out_dim = self.query_projection_size + 2 * self.kv_projection_size
        self.linear_qkv = build_module(
            submodules.linear_qkv, config.hidden_size, self.linear_qkv_out_dim, config=config, init_method=config.init_method,
            gather_output=False, bias=config.add_bias_linear or config.add_qkv_bias, skip_bias_add=False, is_expert=False,
            tp_comm_buffer_name="qkv", tp_group=self.tp_group,
        )
        d = self.hidden_size_per_attention_head
        self.q_layernorm = (
            build_module(submodules.q_layernorm, hidden_size=d, config=config, eps=config.layernorm_epsilon)
            if submodules.q_layernorm is not None else None
        )
        self.k_layernorm = (
            build_module(submodules.k_layernorm, hidden_size=d, config=config, eps=config.layernorm_epsilon)
            if submodules.k_layernorm is not None else None
        )

    def get_query_key_value_tensors(self, hidden_states, key_value_states=None):
        mixed, _ = self.linear_qkv(hidden_states)  # [sq, b, g*(r+2)*d]
        sq, b = mixed.shape[:2]
        g = self.num_query_groups_per_partition
        d = self.hidden_size_per_attention_head
        r = self.num_attention_heads_per_partition // g
        mixed = mixed.view(sq, b, g, (r + 2) * d)
        q, k, v = torch.split(mixed, [r * d, d, d], dim=3)
        q = q.reshape(sq, b, g * r, d)
        if self.q_layernorm is not None:
            q = self.q_layernorm(q)
        if self.k_layernorm is not None:
            k = self.k_layernorm(k)
        return q, k, v


class CrossAttention(Attention):
    def __init__(self, config: TransformerConfig, submodules: CrossAttentionSubmodules, layer_number: int,
                 attn_mask_type=AttnMaskType.padding, cp_comm_type: Optional[str] = None, pg_collection=None):
        super().__init__(config, submodules, layer_number, attn_mask_type, "cross", cp_comm_type, pg_collection)
        if config.num_query_groups != config.num_attention_heads:
            raise ValueError("group query attention is not supported in cross attention")
        self.linear_q = build_module(
            submodules.linear_q, config.hidden_size, self.query_projection_size, config=config, init_method=config.init_method,
            gather_output=False, bias=config.add_bias_linear, skip_bias_add=False, is_expert=False, tp_group=self.tp_group,
        )
        self.linear_kv = build_module(
            submodules.linear_kv, config.hidden_size, 2 * self.kv_projection_size, config=config, init_method=config.init_method,
            gather_output=False, bias=config.add_bias_linear, skip_bias_add=False, is_expert=False, tp_group=self.tp_group,
        )

    def get_query_key_value_tensors(self, hidden_states, key_value_states):
        kv, _ = self.linear_kv(key_value_states)
        d = self.hidden_size_per_attention_head
        kv = kv.view(*kv.shape[:2], self.num_attention_heads_per_partition, 2 * d)
        k, v = torch.split(kv, d, dim=3)
        q, _ = self.linear_q(hidden_states)
        q = q.view(*q.shape[:2], self.num_attention_heads_per_partition, d)
        return q, k, v
