"""BERT (reference ``models/bert/bert_model.py``): embedding(+token types) → bidirectional TransformerBlock →
masked-LM head (tied output layer) + optional NSP/SOP binary head on the pooled ``[CLS]`` state."""
from __future__ import annotations

from typing import Literal, Optional

import torch
from torch import Tensor

from ... import tensor_parallel
from ...enums import ModelType
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig
from ..common.embeddings.language_model_embedding import LanguageModelEmbedding
from ..common.embeddings.rotary_pos_embedding import RotaryEmbedding
from ..common.language_module.language_module import LanguageModule
from .bert_lm_head import BertLMHead
from .pooler import Pooler


def bert_extended_attention_mask(attention_mask: Tensor) -> Tensor:
    """[b, s] 1=keep → [b, 1, s, s] bool, True = masked out."""
    m = attention_mask.unsqueeze(1) * attention_mask.unsqueeze(2)  # [b, s, s]
    return (m.unsqueeze(1) < 0.5)


def bert_position_ids(token_ids: Tensor) -> Tensor:
    s = token_ids.size(1)
    return torch.arange(s, dtype=torch.long, device=token_ids.device).unsqueeze(0).expand_as(token_ids)


class BertModel(LanguageModule):
    def __init__(self, config: TransformerConfig, num_tokentypes: int, transformer_layer_spec: ModuleSpec, vocab_size: int,
                 max_sequence_length: int, pre_process: bool = True, post_process: bool = True, fp16_lm_cross_entropy: bool = False,
                 parallel_output: bool = True, share_embeddings_and_output_weights: bool = False,
                 position_embedding_type: Literal["learned_absolute", "rope"] = "learned_absolute", rotary_percent: float = 1.0,
                 seq_len_interpolation_factor: Optional[float] = None, add_binary_head: bool = True, return_embeddings: bool = False,
                 pg_collection=None, vp_stage: Optional[int] = None):
        super().__init__(config=config, pg_collection=pg_collection)
        if return_embeddings:
            assert post_process and add_binary_head
        self.num_tokentypes, self.vocab_size, self.max_sequence_length = num_tokentypes, vocab_size, max_sequence_length
        self.pre_process, self.post_process = pre_process, post_process
        self.fp16_lm_cross_entropy, self.parallel_output = fp16_lm_cross_entropy, parallel_output
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        self.add_binary_head, self.return_embeddings = add_binary_head, return_embeddings
        self.model_type = ModelType.encoder_or_decoder
        self.vp_stage = vp_stage
        if pre_process:
            self.embedding = LanguageModelEmbedding(config=config, vocab_size=vocab_size, max_sequence_length=max_sequence_length,
                                                    position_embedding_type=position_embedding_type, num_tokentypes=num_tokentypes)
        if position_embedding_type == "rope":
            self.rotary_pos_emb = RotaryEmbedding(kv_channels=config.kv_channels, rotary_percent=rotary_percent,
                                                  rotary_interleaved=config.rotary_interleaved, seq_len_interpolation_factor=seq_len_interpolation_factor,
                                                  use_cpu_initialization=config.use_cpu_initialization)
        self.encoder = TransformerBlock(config=config, spec=transformer_layer_spec, pre_process=pre_process, post_process=post_process,
                                        pg_collection=pg_collection, vp_stage=vp_stage)
        if post_process:
            self.lm_head = BertLMHead(config.hidden_size, config)
            self.output_layer = tensor_parallel.ColumnParallelLinear(
                config.hidden_size, vocab_size, config=config, init_method=config.init_method, bias=True, skip_bias_add=False,
                gather_output=not parallel_output, skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights,
            )
            self.binary_head = None
            if add_binary_head:
                dev = self.lm_head.dense.weight.device
                self.binary_head = torch.nn.Linear(config.hidden_size, 2, device=dev, dtype=config.params_dtype)
                if config.perform_initialization:
                    config.init_method(self.binary_head.weight)
                    self.binary_head.bias.data.zero_()
                self.pooler = Pooler(config.hidden_size, config.init_method, config, config.sequence_parallel)
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, input_tensor):
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        assert len(input_tensor) == 1
        self.encoder.set_input_tensor(input_tensor[0])

    def forward(self, input_ids: Tensor, attention_mask: Tensor, tokentype_ids: Tensor = None, lm_labels: Tensor = None, inference_context=None):
        """``attention_mask`` is the [b, s] padding mask (1 = real token).  Returns (lm loss [b,s] | logits, binary logits)."""
        ext_mask = bert_extended_attention_mask(attention_mask)
        if self.pre_process:
            enc_in = self.embedding(input_ids=input_ids, position_ids=bert_position_ids(input_ids), tokentype_ids=tokentype_ids)
        else:
            enc_in = None
        rotary = None
        if self.position_embedding_type == "rope":
            n = self.rotary_pos_emb.get_rotary_seq_len(inference_context, self.encoder, enc_in, self.config)
            rotary = self.rotary_pos_emb(n)
        hidden = self.encoder(hidden_states=enc_in, attention_mask=ext_mask, inference_context=inference_context, rotary_pos_emb=rotary)
        if not self.post_process:
            return hidden
        binary_logits = None
        if self.add_binary_head:
            pooled = self.pooler(hidden, 0)
            if self.return_embeddings:
                emb = torch.transpose(hidden, 0, 1)
                masks = torch.sum(attention_mask, dim=1)
                out = torch.zeros(emb.shape[0], emb.shape[2], dtype=torch.float32, device=emb.device)
                for i, (e, m) in enumerate(zip(emb, masks)):
                    out[i] = torch.mean(e[1 : m - 1], dim=0)
                return out
            binary_logits = self.binary_head(pooled)
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(self.lm_head(hidden), weight=w)
        if lm_labels is None:
            return logits.transpose(0, 1).contiguous(), binary_logits
        return self.compute_language_model_loss(lm_labels, logits), binary_logits
