"""Hybrid Mamba language model (reference ``models/mamba/mamba_model.py``): embedding → MambaStack → output layer → loss."""
from __future__ import annotations

from typing import Literal, Optional

from torch import Tensor

from ... import tensor_parallel
from ...enums import ModelType
from ...transformer.spec_utils import ModuleSpec, build_module
from ...transformer.transformer_config import TransformerConfig
from ..common.embeddings.language_model_embedding import LanguageModelEmbedding
from ..common.embeddings.rotary_pos_embedding import RotaryEmbedding
from ..common.language_module.language_module import LanguageModule


class MambaModel(LanguageModule):
    def __init__(self, config: TransformerConfig, mamba_stack_spec: ModuleSpec, vocab_size: int, max_sequence_length: int, pre_process: bool = True,
                 hybrid_attention_ratio: float = 0.0, hybrid_mlp_ratio: float = 0.0, hybrid_override_pattern: Optional[str] = None,
                 post_process: bool = True, fp16_lm_cross_entropy: bool = False, parallel_output: bool = True,
                 share_embeddings_and_output_weights: bool = False, position_embedding_type: Literal["learned_absolute", "rope", "none"] = "none",
                 rotary_percent: float = 1.0, rotary_base: int = 10000, seq_len_interpolation_factor: Optional[float] = None, pg_collection=None, vp_stage=None):
        super().__init__(config=config, pg_collection=pg_collection)
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.pre_process, self.post_process = pre_process, post_process
        self.parallel_output, self.share_embeddings_and_output_weights = parallel_output, share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        self.model_type = ModelType.encoder_or_decoder
        self.vp_stage = vp_stage
        if pre_process:
            self.embedding = LanguageModelEmbedding(config=config, vocab_size=vocab_size, max_sequence_length=max_sequence_length,
                                                    position_embedding_type=position_embedding_type)
        if position_embedding_type == "rope":
            self.rotary_pos_emb = RotaryEmbedding(kv_channels=config.kv_channels, rotary_percent=rotary_percent, rotary_base=rotary_base,
                                                  seq_len_interpolation_factor=seq_len_interpolation_factor, use_cpu_initialization=config.use_cpu_initialization)
        self.decoder = build_module(mamba_stack_spec, config, pre_process=pre_process, hybrid_attention_ratio=hybrid_attention_ratio,
                                    hybrid_mlp_ratio=hybrid_mlp_ratio, hybrid_override_pattern=hybrid_override_pattern, post_process=post_process,
                                    pg_collection=pg_collection, vp_stage=vp_stage)
        if post_process:
            self.output_layer = tensor_parallel.ColumnParallelLinear(
                config.hidden_size, vocab_size, config=config, init_method=config.init_method, bias=False, skip_bias_add=False,
                gather_output=not parallel_output, skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights,
            )
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, input_tensor):
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        self.decoder.set_input_tensor(input_tensor[0])

    def forward(self, input_ids: Tensor, position_ids: Tensor, attention_mask: Tensor, decoder_input: Tensor = None, labels: Tensor = None,
                inference_context=None, runtime_gather_output: Optional[bool] = None, *, inference_params=None, loss_mask=None, **_):
        inference_context = inference_context or inference_params
        if decoder_input is None and self.pre_process:
            decoder_input = self.embedding(input_ids=input_ids, position_ids=position_ids)
        rotary = None
        if self.position_embedding_type == "rope":
            n = self.rotary_pos_emb.get_rotary_seq_len(inference_context, self.decoder, decoder_input, self.config)
            rotary = self.rotary_pos_emb(n)
        hidden = self.decoder(hidden_states=decoder_input, attention_mask=attention_mask, inference_context=inference_context, rotary_pos_emb=rotary)
        if not self.post_process:
            return hidden
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(hidden, weight=w, runtime_gather_output=runtime_gather_output)
        if labels is None:
            return logits.transpose(0, 1).contiguous()
        return self.compute_language_model_loss(labels, logits)
