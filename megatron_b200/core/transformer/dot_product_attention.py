"""Core attention (reference ``transformer/dot_product_attention.py:142``).

The reference's local path materialises the ``[b*np, sq, sk]`` score matrix with
``baddbmm`` + softmax + ``bmm``; here the default is the fused sm_100a flash
attention kernel (``ops.flash_attention``), with the unfused math kept as the
CPU / arbitrary-mask path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from ... import ops
from ..enums import AttnMaskType
from ..utils import divide, get_pg_size, get_tensor_model_parallel_group_if_none
from ..tensor_parallel.random import get_cuda_rng_tracker
from .module import MegatronModule
from .transformer_config import TransformerConfig


class DotProductAttention(MegatronModule):
    def __init__(self, config: TransformerConfig, layer_number: int, attn_mask_type: AttnMaskType, attention_type: str,
                 attention_dropout: Optional[float] = None, softmax_scale: Optional[float] = None, cp_comm_type: Optional[str] = None,
                 pg_collection=None, k_channels=None, v_channels=None, **kwargs):
        super().__init__(config)
        self.layer_number = max(1, layer_number)
        self.attn_mask_type = attn_mask_type
        self.attention_type = attention_type
        tp_group = pg_collection.tp if pg_collection is not None and getattr(pg_collection, "tp", None) is not None else get_tensor_model_parallel_group_if_none(None)
        ws = get_pg_size(tp_group)
        kv = config.kv_channels
        proj = kv * config.num_attention_heads
        self.hidden_size_per_partition = divide(proj, ws)
        self.hidden_size_per_attention_head = divide(proj, config.num_attention_heads)
        self.num_attention_heads_per_partition = divide(config.num_attention_heads, ws)
        self.num_query_groups_per_partition = max(1, config.num_query_groups // ws)
        coeff = None
        if softmax_scale is None:
            self.softmax_scale = 1.0 / math.sqrt(k_channels or self.hidden_size_per_attention_head)
        else:
            self.softmax_scale = softmax_scale
        if config.apply_query_key_layer_scaling:
            coeff = self.layer_number
            self.softmax_scale /= coeff
        self.coeff = coeff
        self.dropout_p = config.attention_dropout if attention_dropout is None else attention_dropout
        self.attention_dropout = torch.nn.Dropout(self.dropout_p)
        if config.context_parallel_size > 1:
            from ...parallel.context_parallel import RingAttention

            self.cp = RingAttention(config, cp_comm_type or "p2p", pg_collection)
        else:
            self.cp = None

    def forward(self, query, key, value, attention_mask, attn_mask_type: AttnMaskType = None, attention_bias=None, packed_seq_params=None):
        assert packed_seq_params is None or self.cp is None, "packed sequences + context parallel not supported"
        mask_type = attn_mask_type or self.attn_mask_type
        causal = mask_type in (AttnMaskType.causal, AttnMaskType.padding_causal, AttnMaskType.causal_bottom_right)
        sq, b, hq, d = query.shape
        if self.cp is not None:
            ctx = self.cp(query, key, value, causal, self.softmax_scale)
            return ctx.reshape(sq, b, -1)
        fusable = (
            (attention_mask is None or causal)
            and attention_bias is None
            and (self.dropout_p == 0.0 or not self.training)
        )
        cu = None
        if packed_seq_params is not None:
            # THD: the token dim holds several packed sequences ([t, 1, h, d]); attention must not cross their boundaries (reference: TE's thd kernels
            # driven by cu_seqlens, extensions/transformer_engine.py:1460-1530).  Ours: a band mask inside the native kernels / a block-diagonal mask on CPU.
            cu = packed_seq_params.cu_seqlens_q_padded if getattr(packed_seq_params, "cu_seqlens_q_padded", None) is not None else packed_seq_params.cu_seqlens_q
            assert causal and b == 1 and key.shape[0] == sq, "packed sequences need causal self-attention in THD layout [t, 1, h, d]"
            if not fusable:
                # attention dropout / bias: the unfused path with the block-diagonal causal mask spelled out (O(t²) memory — the kernels cover the dropout-free case)
                pos = torch.arange(sq, device=query.device)
                cul = cu.to(device=query.device, dtype=torch.long)
                sid = torch.bucketize(pos, cul[1:], right=True).clamp_(max=cul.numel() - 2)
                blocked = (sid[:, None] != sid[None, :]) | (pos[None, :] > pos[:, None])
                return self._unfused(query, key, value, blocked.view(1, 1, sq, sq), False, attention_bias)
        if fusable:
            ctx = ops.flash_attention(query, key, value, causal=causal, scale=self.softmax_scale, window=self.config.window_size, cu_seqlens=cu)
            return ctx.reshape(sq, b, -1)
        return self._unfused(query, key, value, attention_mask, causal, attention_bias)

    def _unfused(self, query, key, value, attention_mask, causal, attention_bias):
        sq, b, hq, d = query.shape
        sk, hk = key.shape[0], key.shape[2]
        if hq != hk:
            key = key.repeat_interleave(hq // hk, dim=2)
            value = value.repeat_interleave(hq // hk, dim=2)
        q = query.permute(1, 2, 0, 3).reshape(b * hq, sq, d)
        k = key.permute(1, 2, 0, 3).reshape(b * hq, sk, d)
        scores = torch.bmm(q, k.transpose(1, 2)).view(b, hq, sq, sk)
        if attention_bias is not None:
            scores = scores + attention_bias
        mask = attention_mask
        if mask is not None and mask.dtype != torch.bool:
            mask = mask.bool()
        probs = ops.ref.scaled_masked_softmax(scores, mask, self.softmax_scale, causal=causal and mask is None)
        if self.dropout_p > 0 and self.training:
            if self.config.sequence_parallel:
                probs = self.attention_dropout(probs)
            else:
                with get_cuda_rng_tracker().fork():
                    probs = self.attention_dropout(probs)
        v = value.permute(1, 2, 0, 3).reshape(b * hq, sk, value.shape[-1])
        ctx = torch.bmm(probs.view(b * hq, sq, sk).to(v.dtype), v).view(b, hq, sq, -1)
        return ctx.permute(2, 0, 1, 3).reshape(sq, b, -1)
