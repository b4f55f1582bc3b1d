"""Token (+ learned position, + token-type) embeddings
(reference ``models/common/embeddings/language_model_embedding.py``)."""
from __future__ import annotations


import torch

from megatron_b200.core import tensor_parallel
from megatron_b200.core.transformer.module import MegatronModule
from megatron_b200.core.transformer.transformer_config import TransformerConfig
from megatron_b200.core.utils import get_tensor_model_parallel_group_if_none


class LanguageModelEmbedding(MegatronModule):
    def __init__(self, config: TransformerConfig, vocab_size: int, max_sequence_length: int, position_embedding_type: str = "learned_absolute",
                 num_tokentypes: int = 0, scatter_to_sequence_parallel: bool = True, tp_group=None):
        super().__init__(config)
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.add_position_embedding = position_embedding_type == "learned_absolute"
        self.num_tokentypes = num_tokentypes
        self.scatter_to_sequence_parallel = scatter_to_sequence_parallel
        self.tp_group = get_tensor_model_parallel_group_if_none(tp_group)
        self.reduce_scatter_embeddings = (
            not self.add_position_embedding and num_tokentypes <= 0 and config.sequence_parallel and scatter_to_sequence_parallel
        )
        self.word_embeddings = tensor_parallel.VocabParallelEmbedding(
            vocab_size, config.hidden_size, init_method=config.embedding_init_method,
            reduce_scatter_embeddings=self.reduce_scatter_embeddings, config=config, tp_group=self.tp_group,
        )
        if self.add_position_embedding:
            self.position_embeddings = torch.nn.Embedding(max_sequence_length, config.hidden_size)
            if config.perform_initialization:
                config.embedding_init_method(self.position_embeddings.weight)
            self.position_embeddings.to(dtype=config.params_dtype, device=self.word_embeddings.weight.device)
        if num_tokentypes > 0:
            self.tokentype_embeddings = torch.nn.Embedding(num_tokentypes, config.hidden_size)
            if config.perform_initialization:
                config.init_method(self.tokentype_embeddings.weight)
            self.tokentype_embeddings.to(dtype=config.params_dtype, device=self.word_embeddings.weight.device)
        else:
            self.tokentype_embeddings = None
        self.embedding_dropout = torch.nn.Dropout(config.hidden_dropout)

    def zero_parameters(self):
        self.word_embeddings.weight.data.fill_(0)
        self.word_embeddings.weight.shared = True
        if self.add_position_embedding:
            self.position_embeddings.weight.data.fill_(0)
            self.position_embeddings.weight.shared = True

    def forward(self, input_ids, position_ids, tokentype_ids=None):
        emb = self.word_embeddings(input_ids)
        if self.add_position_embedding:
            emb = emb + self.position_embeddings(position_ids)
        if not self.reduce_scatter_embeddings:
            emb = emb.transpose(0, 1).contiguous()  # [b, s, h] → [s, b, h]
        if tokentype_ids is not None:
            assert self.tokentype_embeddings is not None
            emb = emb + self.tokentype_embeddings(tokentype_ids).permute(1, 0, 2)
        if getattr(self.config, "use_mup", False) and self.config.mup_embedding_mult != 1.0:
            emb = emb * self.config.mup_embedding_mult
        if self.config.fp32_residual_connection:
            emb = emb.float()
        if self.config.sequence_parallel:
            if not self.reduce_scatter_embeddings and self.scatter_to_sequence_parallel:
                emb = tensor_parallel.scatter_to_sequence_parallel_region(emb, group=self.tp_group)
            if self.config.clone_scatter_output_in_embedding and self.scatter_to_sequence_parallel:
                emb = emb.clone()
            with tensor_parallel.get_cuda_rng_tracker().fork():
                emb = self.embedding_dropout(emb)
        else:
            emb = self.embedding_dropout(emb)
        return emb
