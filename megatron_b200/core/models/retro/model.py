"""Retro — retrieval-enhanced transformer (reference ``models/retro/``: ``config.py``, ``encoder_attention.py``, ``decoder_attention.py``, ``model.py``).

A GPT decoder in which selected layers carry *chunked cross-attention* (CCA) to retrieved neighbours.  The sequence of ``n = l·m`` tokens is cut into
``l`` chunks of ``m`` tokens; for every chunk ``k`` neighbours of ``r`` tokens are retrieved offline.  A small bidirectional encoder encodes the
neighbours (conditioned, by cross-attention, on the decoder state of the chunk they were retrieved for), and decoder tokens attend to the encoded
neighbours of the PREVIOUS-or-current chunk with a shift of ``m - 1`` positions, which keeps the model autoregressive:

    token at position  u·m + (m-1) + j   (j = 0..m-1)   attends to   E_u = encode(neighbours of chunk u)

so the first token that can see the neighbours of chunk ``u`` is the LAST token of chunk ``u``.

Layout on B200: the neighbour axis is folded into the batch (``[r, b·l·k, h]``), so both the encoder and the CCA are ordinary dense attention calls
with large batch — no ragged kernels are needed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn.functional as F

from ...transformer.attention import CrossAttention, CrossAttentionSubmodules
from ...transformer.enums import AttnMaskType
from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec, build_module
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig
from ..gpt.gpt_model import GPTModel


@dataclass
class RetroConfig(TransformerConfig):
    retro_chunk_length: int = 64
    retro_num_neighbors: int = 2
    retro_retrieved_length: int = 128              # tokens per neighbour (neighbour chunk + its continuation)
    retro_encoder_num_layers: int = 2
    retro_decoder_cross_attention_layers: Optional[List[int]] = None      # 1-based; default: 6, 9, 12, ... (every third layer from the sixth)
    retro_encoder_hidden_dropout: float = 0.1
    retro_encoder_attention_dropout: float = 0.1

    def __post_init__(self):
        super().__post_init__()
        if self.retro_decoder_cross_attention_layers is None:
            start = 6 if self.num_layers >= 6 else 1
            self.retro_decoder_cross_attention_layers = list(range(start, self.num_layers + 1, 3))


def chunked_cross_attention(attn: CrossAttention, hidden: torch.Tensor, encoded: torch.Tensor, chunk_length: int) -> torch.Tensor:
    """hidden [n, b, h] (n = l·m); encoded [k·r, b·l, h] (neighbours of chunk u of sample i at batch index i·l + u) → [n, b, h] CCA output
    (zeros for the first m-1 positions, which have nothing to attend to)."""
    n, b, h = hidden.shape
    m = chunk_length
    l = n // m
    assert l * m == n, "sequence length must be a multiple of the retro chunk length"
    shifted = F.pad(hidden[m - 1 :], (0, 0, 0, 0, 0, m - 1))                          # attending positions, padded to l full chunks
    q = shifted.view(l, m, b, h).permute(1, 2, 0, 3).reshape(m, b * l, h)              # [m, b·l, h]
    out, bias = attn(q, None, key_value_states=encoded)
    if bias is not None:
        out = out + bias
    out = out.view(m, b, l, h).permute(2, 0, 1, 3).reshape(n, b, h)
    return F.pad(out, (0, 0, 0, 0, m - 1, 0))[:n]                                      # shift back: position p gets what was computed at p - (m-1)


class RetroEncoder(MegatronModule):
    """Bidirectional transformer over each neighbour; its layers cross-attend to the decoder state of the chunk the neighbour belongs to."""

    def __init__(self, config: RetroConfig, layer_spec: ModuleSpec, cross_spec: ModuleSpec):
        super().__init__(config)
        import copy

        ecfg = copy.copy(config)
        ecfg.num_layers = config.retro_encoder_num_layers
        ecfg.hidden_dropout, ecfg.attention_dropout = config.retro_encoder_hidden_dropout, config.retro_encoder_attention_dropout
        ecfg.pipeline_model_parallel_size = 1
        self.block = TransformerBlock(config=ecfg, spec=layer_spec, pre_process=True, post_process=True)
        self.cross = build_module(cross_spec, config=ecfg, layer_number=1, attn_mask_type=AttnMaskType.padding)
        self.cross_norm = torch.nn.LayerNorm(config.hidden_size, eps=config.layernorm_epsilon, dtype=config.params_dtype)

    def forward(self, neighbour_emb: torch.Tensor, chunk_states: torch.Tensor) -> torch.Tensor:
        """neighbour_emb [r, b·l·k, h]; chunk_states [m, b·l, h] (decoder hidden of each chunk) → [r, b·l·k, h]."""
        k = neighbour_emb.shape[1] // chunk_states.shape[1]
        ctx = chunk_states.repeat_interleave(k, dim=1)
        x, bias = self.cross(self.cross_norm(neighbour_emb), None, key_value_states=ctx)
        x = neighbour_emb + (x if bias is None else x + bias)
        return self.block(x, None)


class RetroModel(GPTModel):
    """``forward(input_ids, position_ids, attention_mask, context_input_ids, context_position_ids, labels=...)``.
    ``context_input_ids`` : ``[b, l, k, r]`` retrieved neighbour tokens for every chunk."""

    def __init__(self, config: RetroConfig, transformer_layer_spec: ModuleSpec, vocab_size: int, max_sequence_length: int, encoder_layer_spec: Optional[ModuleSpec] = None,
                 cross_attention_spec: Optional[ModuleSpec] = None, **kwargs):
        super().__init__(config, transformer_layer_spec, vocab_size, max_sequence_length, **kwargs)
        from ...tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
        from ...transformer.dot_product_attention import DotProductAttention

        cross_attention_spec = cross_attention_spec or ModuleSpec(module=CrossAttention, submodules=CrossAttentionSubmodules(
            linear_q=ColumnParallelLinear, linear_kv=ColumnParallelLinear, core_attention=DotProductAttention, linear_proj=RowParallelLinear))
        self.retro_layers = sorted(config.retro_decoder_cross_attention_layers)
        self.cca = torch.nn.ModuleDict({str(i): build_module(cross_attention_spec, config=config, layer_number=i, attn_mask_type=AttnMaskType.padding)
                                        for i in self.retro_layers})
        self.cca_norm = torch.nn.ModuleDict({str(i): torch.nn.LayerNorm(config.hidden_size, eps=config.layernorm_epsilon, dtype=config.params_dtype)
                                             for i in self.retro_layers})
        self.encoder = RetroEncoder(config, encoder_layer_spec or transformer_layer_spec, cross_attention_spec)

    def _encode_neighbours(self, context_input_ids, context_position_ids, hidden):
        b, l, k, r = context_input_ids.shape
        m = self.config.retro_chunk_length
        ids = context_input_ids.reshape(b * l * k, r)
        pos = context_position_ids.reshape(b * l * k, r) if context_position_ids is not None else torch.arange(r, device=ids.device)[None].expand_as(ids)
        emb = self.embedding(input_ids=ids, position_ids=pos)                              # [r, b·l·k, h]
        n = hidden.shape[0]
        chunks = hidden.view(l, m, b, -1).permute(1, 2, 0, 3).reshape(m, b * l, -1)        # decoder state per chunk
        enc = self.encoder(emb, chunks)                                                    # [r, b·l·k, h]
        return enc.view(r, b * l, k, -1).permute(2, 0, 1, 3).reshape(k * r, b * l, -1)     # neighbours of a chunk side by side along the key axis

    def forward(self, input_ids, position_ids, attention_mask=None, context_input_ids=None, context_position_ids=None, labels=None, loss_mask=None, **kw):
        decoder_input, rotary = self._preprocess(input_ids, position_ids)
        h = decoder_input
        encoded = None
        for layer in self.decoder.layers:
            h, _ = layer(h, attention_mask=attention_mask, rotary_pos_emb=rotary)
            ln = layer.layer_number
            if context_input_ids is not None and ln in self.retro_layers:
                if encoded is None:          # encoded once, at the first retrieval layer, from that layer's decoder state (as in the paper)
                    encoded = self._encode_neighbours(context_input_ids, context_position_ids, h)
                h = h + chunked_cross_attention(self.cca[str(ln)], self.cca_norm[str(ln)](h), encoded, self.config.retro_chunk_length)
        if self.decoder.final_layernorm is not None:
            h = self.decoder.final_layernorm(h)
        return self._postprocess(h, input_ids, position_ids, labels, rotary, loss_mask, attention_mask, None, None, None)
