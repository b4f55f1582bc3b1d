"""T5 encoder-decoder (reference ``models/T5/t5_model.py``).

encoder: embedding → bidirectional block (padding mask);  decoder: shared embedding → causal self-attn + cross-attn on the
encoder output → LM head (dense-free: tied output layer + bias).  Masks follow the reference convention
(``True`` = masked out)."""
from __future__ import annotations

from typing import Literal, Optional

import torch
from torch import Tensor

from ... import tensor_parallel
from ...enums import ModelType
from ...transformer.module import MegatronModule
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig
from ..common.embeddings.language_model_embedding import LanguageModelEmbedding
from ..common.embeddings.rotary_pos_embedding import RotaryEmbedding
from ..common.language_module.language_module import LanguageModule


class T5LMHead(MegatronModule):
    def __init__(self, config: TransformerConfig, parallel_output: bool, vocab_size: int, pre_process: bool = True, share_embeddings_and_output_weights: bool = False):
        super().__init__(config)
        self.parallel_output = parallel_output
        self.output_layer = tensor_parallel.ColumnParallelLinear(
            config.hidden_size, vocab_size, config=config, init_method=config.init_method, bias=share_embeddings_and_output_weights,
            skip_bias_add=not share_embeddings_and_output_weights, gather_output=not parallel_output,
            skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights,
        )

    def forward(self, hidden_states: Tensor, word_embeddings_weight: Tensor) -> Tensor:
        logits, _ = self.output_layer(hidden_states, weight=word_embeddings_weight)
        return logits


def t5_extended_attention_mask(masks):
    """→ [b, 1, sq, sk] bool with True = masked.  Accepts the reference's calling convention (already-extended boolean masks ``[b, 1, sq, sk]``, True = masked:
    what ``pretrain_t5.py`` builds) and this framework's ``[b, sq, sk]`` keep-masks (1 = attend)."""
    out = []
    for m in masks:
        if m is None:
            out.append(None)
        elif m.dim() == 4:
            out.append(m if m.dtype == torch.bool else m > 0.5)
        else:
            out.append(m.unsqueeze(1) < 0.5)
    return out


def t5_position_ids(token_ids: Tensor) -> Tensor:
    return torch.arange(token_ids.size(1), dtype=torch.long, device=token_ids.device).unsqueeze(0).expand_as(token_ids)


class T5Model(LanguageModule):
    def __init__(self, config: TransformerConfig, encoder_config: TransformerConfig, transformer_encoder_layer_spec, transformer_decoder_layer_spec,
                 vocab_size: int, max_sequence_length: int, pre_process: bool = True, post_process: bool = True, fp16_lm_cross_entropy: bool = False,
                 parallel_output: bool = True, share_embeddings_and_output_weights: bool = False,
                 position_embedding_type: Literal["learned_absolute", "rope", "relative"] = "learned_absolute", rotary_percent: float = 1.0,
                 seq_len_interpolation_factor: Optional[float] = None, add_encoder: bool = True, add_decoder: bool = True, pg_collection=None, vp_stage=None):
        super().__init__(config=config, pg_collection=pg_collection)
        self.encoder_config = encoder_config
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.pre_process, self.post_process = pre_process, post_process
        self.add_encoder, self.add_decoder = add_encoder, add_decoder
        self.fp16_lm_cross_entropy, self.parallel_output = fp16_lm_cross_entropy, parallel_output
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        self.encoder_hidden_state = None
        self.model_type = ModelType.encoder_or_decoder
        self.vp_stage = vp_stage
        if pre_process:
            self.embedding = LanguageModelEmbedding(config=config, vocab_size=vocab_size, max_sequence_length=max_sequence_length,
                                                    position_embedding_type=position_embedding_type)
        if position_embedding_type == "rope":
            self.rotary_pos_emb = RotaryEmbedding(kv_channels=config.kv_channels, rotary_percent=rotary_percent, rotary_interleaved=config.rotary_interleaved,
                                                  seq_len_interpolation_factor=seq_len_interpolation_factor, use_cpu_initialization=config.use_cpu_initialization)
        self.encoder = TransformerBlock(config=encoder_config, spec=transformer_encoder_layer_spec, pre_process=pre_process, post_process=post_process,
                                        pg_collection=pg_collection) if add_encoder else None
        self.decoder = TransformerBlock(config=config, spec=transformer_decoder_layer_spec, pre_process=pre_process, post_process=post_process,
                                        pg_collection=pg_collection) if add_decoder else None
        if post_process:
            self.lm_head = T5LMHead(config, parallel_output, vocab_size, pre_process, share_embeddings_and_output_weights)
            self.output_layer = self.lm_head.output_layer
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, input_tensor):
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        if self.add_encoder and self.add_decoder:
            assert len(input_tensor) == 1
            self.encoder.set_input_tensor(input_tensor[0])
        elif self.add_encoder:
            self.encoder.set_input_tensor(input_tensor[0])
        elif self.add_decoder:
            if len(input_tensor) == 2:
                self.decoder.set_input_tensor(input_tensor[0])
                self.encoder_hidden_state = input_tensor[1]
            else:
                self.decoder.set_input_tensor(None)
                self.encoder_hidden_state = input_tensor[0]

    def forward(self, encoder_input_ids: Tensor, decoder_input_ids: Tensor, encoder_attn_mask: Tensor, decoder_attn_mask: Tensor,
                encoder_decoder_attn_mask: Tensor, lm_labels: Tensor = None, encoder_hidden_states: Tensor = None, output_encoder_hidden_only: bool = False,
                inference_context=None, packed_seq_params=None):
        enc_mask, dec_mask, xattn_mask = t5_extended_attention_mask([encoder_attn_mask, decoder_attn_mask, encoder_decoder_attn_mask])
        rotary = None
        if self.position_embedding_type == "rope":
            rotary = self.rotary_pos_emb(max(encoder_input_ids.size(1), decoder_input_ids.size(1)))
        if encoder_hidden_states is None and self.encoder_hidden_state is not None:
            encoder_hidden_states = self.encoder_hidden_state
        if encoder_hidden_states is None and self.add_encoder:
            enc_in = self.embedding(input_ids=encoder_input_ids, position_ids=t5_position_ids(encoder_input_ids)) if self.pre_process else None
            encoder_hidden_states = self.encoder(hidden_states=enc_in, attention_mask=enc_mask, inference_context=inference_context,
                                                 rotary_pos_emb=None if rotary is None else rotary[: encoder_input_ids.size(1)])
        if not self.add_decoder or output_encoder_hidden_only:
            return encoder_hidden_states
        dec_in = self.embedding(input_ids=decoder_input_ids, position_ids=t5_position_ids(decoder_input_ids)) if self.pre_process else None
        # causal ∧ padding mask for decoder self-attention
        sd = decoder_input_ids.size(1)
        causal = torch.triu(torch.ones(sd, sd, dtype=torch.bool, device=decoder_input_ids.device), diagonal=1)[None, None]
        dec_full = causal if dec_mask is None else (dec_mask | causal)
        hidden = self.decoder(hidden_states=dec_in, attention_mask=dec_full, context=encoder_hidden_states, context_mask=xattn_mask,
                              inference_context=inference_context, rotary_pos_emb=None if rotary is None else rotary[:sd])
        if not self.post_process:
            return hidden
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits = self.lm_head(hidden, word_embeddings_weight=w)
        if lm_labels is None:
            return logits.transpose(0, 1).contiguous()
        return self.compute_language_model_loss(lm_labels, logits)
