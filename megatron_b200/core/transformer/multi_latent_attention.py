"""Multi-latent attention (DeepSeek-V2/V3) — reference ``transformer/multi_latent_attention.py`` (1,502 LoC).

Queries and keys/values are produced through low-rank "latent" bottlenecks::

    h ─ W_dq ─ norm ─ W_uq ─►  q = [q_nope | q_pe]            (per head: qk_head_dim + qk_pos_emb_head_dim)
    h ─ W_dkv ─┬─ norm ─ W_ukv ─► [k_nope | v]                 (per head: qk_head_dim + v_head_dim)
               └─ k_pe (ONE rotary key shared by all heads)

RoPE (YaRN-scaled) touches only the ``*_pe`` slices.  The down projections are replicated over TP (they are
tiny); the up projections are column-parallel over heads, the output projection row-parallel.  Under sequence
parallelism the down projections run on the local ``[s/tp]`` shard (their weight grads are then all-reduced
over TP via the ``sequence_parallel`` tag) and the up projections all-gather inside their fused AG→GEMM op.

For inference ``cache_mla_latents`` stores the compressed ``[kv_latent | k_pe]`` (``kv_lora_rank +
qk_pos_emb_head_dim`` values per token instead of ``2·n·d``) and re-expands it on the fly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import torch

from ... import ops
from ..enums import AttnMaskType
from ..models.common.embeddings.rotary_pos_embedding import RotaryEmbedding
from ..models.common.embeddings.yarn_rotary_pos_embedding import YarnRotaryEmbedding, yarn_get_mscale
from ..tensor_parallel.mappings import gather_from_sequence_parallel_region
from ..utils import divide, get_pg_size, get_tensor_model_parallel_group_if_none
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import MLATransformerConfig


@dataclass
class MLASelfAttentionSubmodules:
    linear_q_proj: Union[ModuleSpec, type] = None
    linear_q_down_proj: Union[ModuleSpec, type] = None
    linear_q_up_proj: Union[ModuleSpec, type] = None
    linear_kv_down_proj: Union[ModuleSpec, type] = None
    linear_kv_up_proj: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None
    q_layernorm: Union[ModuleSpec, type] = None
    kv_layernorm: Union[ModuleSpec, type] = None


class ReplicatedLinear(torch.nn.Module):
    """``y = x Wᵀ`` with the weight replicated on every TP rank (MLA down projections)."""

    def __init__(self, input_size: int, output_size: int, *, config, init_method=None, bias: bool = False, **_):
        super().__init__()
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight = torch.nn.Parameter(torch.empty(output_size, input_size, dtype=config.params_dtype, device=dev))
        if config.perform_initialization:
            (init_method or config.init_method)(self.weight)
        # sharded-sequence input ⇒ partial weight grads on each TP rank ⇒ all-reduce them like norm weights
        setattr(self.weight, "sequence_parallel", bool(config.sequence_parallel))
        self.bias = torch.nn.Parameter(torch.zeros(output_size, dtype=config.params_dtype, device=dev)) if bias else None
        if self.bias is not None:
            setattr(self.bias, "sequence_parallel", bool(config.sequence_parallel))

    def forward(self, x):
        from ..tensor_parallel.layers import linear_with_grad_accumulation_and_async_allreduce

        y = linear_with_grad_accumulation_and_async_allreduce(x, self.weight, self.bias, False, False, False)
        return y, None

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from .utils import make_sharded_tensors_for_checkpoint

        return make_sharded_tensors_for_checkpoint(self.state_dict(prefix="", keep_vars=True), prefix, {}, sharded_offsets)


def _apply_rope(t, angles, mscale: float, interleaved: bool):
    """MLA rotary: optional de-interleave (DeepSeek stores pairs adjacent), rotate, scale by YaRN mscale."""
    if interleaved:
        t = torch.cat([t[..., 0::2], t[..., 1::2]], dim=-1)
    return ops.apply_rope(t, angles, False, mscale)


class MLASelfAttention(MegatronModule):
    def __init__(self, config: MLATransformerConfig, submodules: MLASelfAttentionSubmodules, layer_number: int,
                 attn_mask_type=AttnMaskType.causal, cp_comm_type: Optional[str] = None, pg_collection=None):
        super().__init__(config)
        self.layer_number, self.attn_mask_type = layer_number, attn_mask_type
        self.tp_group = pg_collection.tp if pg_collection is not None and getattr(pg_collection, "tp", None) is not None else get_tensor_model_parallel_group_if_none(None)
        ws = get_pg_size(self.tp_group)
        c = config
        self.n_heads = c.num_attention_heads
        self.n_local = divide(self.n_heads, ws)
        self.q_head_dim = c.qk_head_dim + c.qk_pos_emb_head_dim
        self.softmax_mscale = yarn_get_mscale(c.rotary_scaling_factor, c.mscale_all_dim) if c.rope_type == "yarn" else 1.0
        self.softmax_scale = self.softmax_mscale * self.softmax_mscale / math.sqrt(self.q_head_dim)
        if c.rope_type == "yarn":
            self.rotary_pos_emb = YarnRotaryEmbedding(
                c.qk_pos_emb_head_dim, rotary_base=c.rotary_base, scaling_factor=c.rotary_scaling_factor,
                original_max_position_embeddings=c.original_max_position_embeddings, beta_fast=c.beta_fast, beta_slow=c.beta_slow,
                mscale=c.mscale, mscale_all_dim=c.mscale_all_dim, use_cpu_initialization=c.use_cpu_initialization,
            )
        else:
            self.rotary_pos_emb = RotaryEmbedding(c.qk_pos_emb_head_dim, rotary_percent=c.rotary_percent, rotary_base=c.rotary_base,
                                                  use_cpu_initialization=c.use_cpu_initialization)
        col = dict(config=c, init_method=c.init_method, gather_output=False, bias=False, skip_bias_add=False, is_expert=False, tp_group=self.tp_group)
        if c.q_lora_rank is None:
            self.linear_q_proj = build_module(submodules.linear_q_proj, c.hidden_size, self.n_heads * self.q_head_dim, **col)
            self.linear_q_down_proj = self.linear_q_up_proj = self.q_layernorm = None
        else:
            self.linear_q_proj = None
            self.linear_q_down_proj = ReplicatedLinear(c.hidden_size, c.q_lora_rank, config=c, init_method=c.init_method)
            self.linear_q_up_proj = build_module(submodules.linear_q_up_proj, c.q_lora_rank, self.n_heads * self.q_head_dim, **col)
            self.q_layernorm = build_module(submodules.q_layernorm, hidden_size=c.q_lora_rank, config=c, eps=c.layernorm_epsilon)
        self.linear_kv_down_proj = ReplicatedLinear(c.hidden_size, c.kv_lora_rank + c.qk_pos_emb_head_dim, config=c, init_method=c.init_method)
        self.linear_kv_up_proj = build_module(submodules.linear_kv_up_proj, c.kv_lora_rank, self.n_heads * (c.qk_head_dim + c.v_head_dim), **col)
        self.kv_layernorm = build_module(submodules.kv_layernorm, hidden_size=c.kv_lora_rank, config=c, eps=c.layernorm_epsilon)
        self.core_attention = build_module(
            submodules.core_attention, config=c, layer_number=layer_number, attn_mask_type=attn_mask_type, attention_type="self",
            softmax_scale=self.softmax_scale, k_channels=self.q_head_dim, v_channels=c.v_head_dim, cp_comm_type=cp_comm_type, pg_collection=pg_collection,
        )
        self.linear_proj = build_module(
            submodules.linear_proj, self.n_heads * c.v_head_dim, c.hidden_size, config=c, init_method=c.output_layer_init_method,
            bias=c.add_bias_linear, input_is_parallel=True, skip_bias_add=True, is_expert=False, tp_group=self.tp_group,
        )
        self.sequence_parallel = c.sequence_parallel and ws > 1
        self.fused_rope = True   # CUDA: in-place rotary on q and the one-kernel kv split (fusions/fused_mla_yarn_rope_apply.py); off → split / rotate / cat in PyTorch

    # ---- projections ------------------------------------------------------------------------------------------
    def _expand_kv(self, kv_latent, k_pe):
        """latent [s,b,r] (normalised) + rotated k_pe [s,b,1,dp] → K [s,b,n,dq], V [s,b,n,dv]."""
        c = self.config
        kv, _ = self.linear_kv_up_proj(kv_latent)
        s, b = kv.shape[:2]
        kv = kv.view(s, b, self.n_local, c.qk_head_dim + c.v_head_dim)
        if kv.is_cuda and self.fused_rope:
            # split + broadcast of the shared rotary key + concatenate in ONE kernel (k_pe is already rotated here); backward sums the rotary gradient over heads
            from ..fusions.fused_mla_yarn_rope_apply import fused_apply_mla_rope_for_kv

            return fused_apply_mla_rope_for_kv(kv, k_pe, None, c.qk_pos_emb_head_dim, c.qk_head_dim, c.v_head_dim)
        k_nope, v = torch.split(kv, [c.qk_head_dim, c.v_head_dim], dim=-1)
        k = torch.cat([k_nope, k_pe.expand(s, b, self.n_local, c.qk_pos_emb_head_dim)], dim=-1)
        return k, v

    def get_query_key_value_tensors(self, hidden_states, inference_context=None):
        c = self.config
        if self.linear_q_proj is not None:
            q, _ = self.linear_q_proj(hidden_states)
        else:
            qc, _ = self.linear_q_down_proj(hidden_states)
            q, _ = self.linear_q_up_proj(self.q_layernorm(qc))
        s, b = q.shape[:2]
        q = q.view(s, b, self.n_local, self.q_head_dim)
        kvc, _ = self.linear_kv_down_proj(hidden_states)
        kv_latent, k_pe = torch.split(kvc, [c.kv_lora_rank, c.qk_pos_emb_head_dim], dim=-1)
        kv_latent = self.kv_layernorm(kv_latent)
        if self.sequence_parallel:
            k_pe = gather_from_sequence_parallel_region(k_pe, group=self.tp_group)
        k_pe = k_pe.unsqueeze(2)  # [s, b, 1, dp]
        # rotary angles for the positions present in this call
        off = inference_context.sequence_len_offset if inference_context is not None else 0
        total = off + s
        emb = self.rotary_pos_emb(total)
        mscale = 1.0
        if isinstance(emb, tuple):
            emb, mscale = emb
        ang = emb[off:total]
        k_pe = _apply_rope(k_pe, ang, mscale, c.rotary_interleaved)
        if q.is_cuda and q.is_contiguous() and self.fused_rope and c.qk_pos_emb_head_dim <= 64:
            # rotate the trailing rotary channels of every head in place: no split / rotate / concatenate round trip over q
            from ..fusions.fused_mla_yarn_rope_apply import fused_apply_mla_rope_for_q

            q = fused_apply_mla_rope_for_q(q, ang, c.qk_head_dim, c.qk_pos_emb_head_dim, mscale, c.rotary_interleaved)
        else:
            q_nope, q_pe = torch.split(q, [c.qk_head_dim, c.qk_pos_emb_head_dim], dim=-1)
            q_pe = _apply_rope(q_pe, ang, mscale, c.rotary_interleaved)
            q = torch.cat([q_nope, q_pe], dim=-1)
        return q, kv_latent, k_pe

    def forward(self, hidden_states, attention_mask, key_value_states=None, inference_context=None, rotary_pos_emb=None,
                rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None, packed_seq_params=None, sequence_len_offset=None,
                *, inference_params=None):
        inference_context = inference_context or inference_params
        c = self.config
        q, kv_latent, k_pe = self.get_query_key_value_tensors(hidden_states, inference_context)
        mask_type = self.attn_mask_type
        if inference_context is not None:
            # latent KV cache: [kv_latent | k_pe] per token (reference cache_mla_latents path)
            kvd = inference_context.key_value_memory_dict
            lat = torch.cat([kv_latent if not self.sequence_parallel else gather_from_sequence_parallel_region(kv_latent, group=self.tp_group), k_pe.squeeze(2)], dim=-1)
            if self.layer_number not in kvd:
                kvd[self.layer_number] = torch.empty(inference_context.max_sequence_length, inference_context.max_batch_size, lat.shape[-1], dtype=lat.dtype, device=lat.device)
            cache = kvd[self.layer_number]
            s0, b0 = inference_context.sequence_len_offset, inference_context.batch_size_offset
            s1, b1 = s0 + lat.shape[0], b0 + lat.shape[1]
            cache[s0:s1, b0:b1] = lat
            full = cache[:s1, b0:b1]
            kv_latent, k_pe = full[..., : c.kv_lora_rank], full[..., c.kv_lora_rank :].unsqueeze(2)
            if s0 > 0 and q.shape[0] == 1:
                mask_type = AttnMaskType.no_mask
            k, v = self._expand_kv_no_sp(kv_latent, k_pe)
        else:
            k, v = self._expand_kv(kv_latent, k_pe)
        # the fused kernel wants equal q/k/v head dims: zero-pad V and drop the padding afterwards
        dv = c.v_head_dim
        if dv != self.q_head_dim:
            v = torch.nn.functional.pad(v, (0, self.q_head_dim - dv))
        out = self.core_attention(q, k, v, attention_mask, attn_mask_type=mask_type, attention_bias=attention_bias, packed_seq_params=packed_seq_params)
        if dv != self.q_head_dim:
            s, b = out.shape[:2]
            out = out.view(s, b, self.n_local, self.q_head_dim)[..., :dv].reshape(s, b, self.n_local * dv)
        return self.linear_proj(out)

    def _expand_kv_no_sp(self, kv_latent, k_pe):
        """Inference: the cached latents are already full-sequence, so bypass the SP all-gather of the up-proj."""
        if not self.sequence_parallel:
            return self._expand_kv(kv_latent, k_pe)
        c = self.config
        kv = torch.nn.functional.linear(kv_latent, self.linear_kv_up_proj.weight)
        s, b = kv.shape[:2]
        kv = kv.view(s, b, self.n_local, c.qk_head_dim + c.v_head_dim)
        if kv.is_cuda and self.fused_rope:
            # split + broadcast of the shared rotary key + concatenate in ONE kernel (k_pe is already rotated here); backward sums the rotary gradient over heads
            from ..fusions.fused_mla_yarn_rope_apply import fused_apply_mla_rope_for_kv

            return fused_apply_mla_rope_for_kv(kv, k_pe, None, c.qk_pos_emb_head_dim, c.qk_head_dim, c.v_head_dim)
        k_nope, v = torch.split(kv, [c.qk_head_dim, c.v_head_dim], dim=-1)
        return torch.cat([k_nope, k_pe.expand(s, b, self.n_local, c.qk_pos_emb_head_dim)], dim=-1), v
