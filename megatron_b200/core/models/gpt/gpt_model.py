"""GPT / Llama-family decoder-only model (reference ``models/gpt/gpt_model.py:52-308``).

embedding → (rotary angles) → TransformerBlock → [MTP] → output_layer → loss.
Pipeline stages own ``pre_process`` (embedding) and/or ``post_process`` (head+loss).
"""
from __future__ import annotations

from typing import Dict, Literal, Optional

from torch import Tensor

from ... import tensor_parallel
from ...dist_checkpointing.mapping import ShardedStateDict
from ...enums import ModelType
from ...packed_seq_params import PackedSeqParams
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig
from ..common.embeddings.language_model_embedding import LanguageModelEmbedding
from ..common.embeddings.rotary_pos_embedding import RotaryEmbedding
from ..common.embeddings.yarn_rotary_pos_embedding import YarnRotaryEmbedding
from ..common.language_module.language_module import LanguageModule


class GPTModel(LanguageModule):
    def __init__(
        self,
        config: TransformerConfig,
        transformer_layer_spec: ModuleSpec,
        vocab_size: int,
        max_sequence_length: int,
        pre_process: bool = True,
        post_process: bool = True,
        fp16_lm_cross_entropy: bool = False,
        logit_dtype=None,
        parallel_output: bool = True,
        share_embeddings_and_output_weights: bool = False,
        position_embedding_type: Literal["learned_absolute", "rope", "mrope", "yarn", "none"] = "learned_absolute",
        rotary_percent: float = 1.0,
        rotary_base: int = 10000,
        rope_scaling: bool = False,
        rope_scaling_factor: float = 8.0,
        scatter_embedding_sequence_parallel: bool = True,
        seq_len_interpolation_factor: Optional[float] = None,
        mtp_block_spec: Optional[ModuleSpec] = None,
        pg_collection=None,
        vp_stage: Optional[int] = None,
    ) -> None:
        super().__init__(config=config, pg_collection=pg_collection)
        self.transformer_layer_spec = transformer_layer_spec
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.pre_process, self.post_process = pre_process, post_process
        self.fp16_lm_cross_entropy = fp16_lm_cross_entropy
        self.logit_dtype = logit_dtype
        self.parallel_output = parallel_output
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        self.vp_stage = vp_stage
        self.model_type = ModelType.encoder_or_decoder
        self.max_position_embeddings = max_sequence_length
        self.rotary_percent = rotary_percent
        self.mtp_process = mtp_block_spec is not None and config.mtp_num_layers

        if self.pre_process or self.mtp_process:
            self.embedding = LanguageModelEmbedding(
                config=config, vocab_size=vocab_size, max_sequence_length=max_sequence_length,
                position_embedding_type=position_embedding_type, scatter_to_sequence_parallel=scatter_embedding_sequence_parallel,
            )
        if position_embedding_type == "rope" and not config.multi_latent_attention:
            self.rotary_pos_emb = RotaryEmbedding(
                kv_channels=config.kv_channels, rotary_percent=rotary_percent, rotary_interleaved=config.rotary_interleaved,
                seq_len_interpolation_factor=seq_len_interpolation_factor, rotary_base=rotary_base, rope_scaling=rope_scaling,
                rope_scaling_factor=rope_scaling_factor, use_cpu_initialization=config.use_cpu_initialization,
            )
        elif position_embedding_type == "yarn":
            self.rotary_pos_emb = YarnRotaryEmbedding(
                kv_channels=config.kv_channels, rotary_percent=rotary_percent, rotary_interleaved=config.rotary_interleaved,
                rotary_base=rotary_base, scaling_factor=rope_scaling_factor, use_cpu_initialization=config.use_cpu_initialization,
            )
        elif position_embedding_type == "mrope":
            raise NotImplementedError("mrope lives in the multimodal models")

        self.decoder = TransformerBlock(
            config=config, spec=transformer_layer_spec, pre_process=pre_process, post_process=post_process,
            pg_collection=pg_collection, vp_stage=vp_stage,
        )
        if self.mtp_process:
            from ...transformer.multi_token_prediction import MultiTokenPredictionBlock

            self.mtp = MultiTokenPredictionBlock(config=config, spec=mtp_block_spec, vp_stage=vp_stage)

        if self.post_process or self.mtp_process:
            if config.defer_embedding_wgrad_compute:
                self.embedding_activation_buffer, self.grad_output_buffer = [], []
            else:
                self.embedding_activation_buffer = self.grad_output_buffer = None
            self.output_layer = tensor_parallel.ColumnParallelLinear(
                config.hidden_size, vocab_size, config=config, bias=False, skip_bias_add=False,
                init_method=config.embedding_init_method if getattr(config, "use_mup", False) and not share_embeddings_and_output_weights else config.init_method,
                gather_output=not parallel_output,
                skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights,
                embedding_activation_buffer=self.embedding_activation_buffer, grad_output_buffer=self.grad_output_buffer,
            )
        if self.pre_process or self.post_process:
            self.setup_embeddings_and_output_layer()
        if getattr(config, "quant_recipe", None) is not None:
            # per-layer precision from module-path globs (core/quantization): which projections run FP8 / MXFP8 / NVFP4, which stay bf16
            from ...quantization import apply_quantization_recipe

            self.quantized_layers = apply_quantization_recipe(self, config.quant_recipe)

    def set_input_tensor(self, input_tensor: Tensor) -> None:
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        assert len(input_tensor) == 1, "input_tensor should only be length 1 for gpt"
        self.decoder.set_input_tensor(input_tensor[0])

    def _preprocess(self, input_ids, position_ids, decoder_input=None, inference_context=None, packed_seq_params=None):
        if decoder_input is not None:
            pass
        elif self.pre_process:
            decoder_input = self.embedding(input_ids=input_ids, position_ids=position_ids)
        else:
            decoder_input = None  # comes from set_input_tensor
        rotary_pos_emb = None
        if self.position_embedding_type in ("rope", "yarn") and not self.config.multi_latent_attention:
            n = self.rotary_pos_emb.get_rotary_seq_len(inference_context, self.decoder, decoder_input, self.config, packed_seq_params)
            rotary_pos_emb = self.rotary_pos_emb(n, packed_seq=packed_seq_params is not None and getattr(packed_seq_params, "qkv_format", "") == "thd")
            if isinstance(rotary_pos_emb, tuple):  # yarn returns (angles, mscale)
                rotary_pos_emb = rotary_pos_emb[0]
        return decoder_input, rotary_pos_emb

    def forward(
        self,
        input_ids: Tensor,
        position_ids: Tensor,
        attention_mask: Tensor,
        decoder_input: Tensor = None,
        labels: Tensor = None,
        inference_context=None,
        packed_seq_params: PackedSeqParams = None,
        extra_block_kwargs: dict = None,
        runtime_gather_output: Optional[bool] = None,
        *,
        inference_params=None,
        loss_mask: Optional[Tensor] = None,
        padding_mask: Optional[Tensor] = None,
    ) -> Tensor:
        """Returns per-token loss ``[b, s]`` when ``labels`` is given, else logits ``[s, b, v/tp]``
        (or the hidden state on non-final pipeline stages)."""
        inference_context = inference_context or inference_params
        decoder_input, rotary_pos_emb = self._preprocess(input_ids, position_ids, decoder_input, inference_context, packed_seq_params)
        hidden_states = self.decoder(
            hidden_states=decoder_input, attention_mask=attention_mask, inference_context=inference_context,
            rotary_pos_emb=rotary_pos_emb, packed_seq_params=packed_seq_params, **(extra_block_kwargs or {}),
        )
        return self._postprocess(hidden_states, input_ids, position_ids, labels, rotary_pos_emb, loss_mask, attention_mask,
                                 packed_seq_params, runtime_gather_output, inference_context)

    def _postprocess(self, hidden_states, input_ids, position_ids, labels, rotary_pos_emb, loss_mask, attention_mask,
                     packed_seq_params, runtime_gather_output, inference_context):
        if not self.post_process:
            return hidden_states
        output_weight = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        if self.mtp_process:
            hidden_states = self.mtp(
                input_ids=input_ids, position_ids=position_ids, labels=labels, loss_mask=loss_mask, hidden_states=hidden_states,
                attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb, embedding=self.embedding, output_layer=self.output_layer,
                output_weight=output_weight, compute_language_model_loss=self.compute_language_model_loss,
            )
        if inference_context is not None and getattr(inference_context, "materialize_only_last_token_logits", False):
            hidden_states = hidden_states[-1:, :, :]
        logits, _ = self.output_layer(hidden_states, weight=output_weight, runtime_gather_output=runtime_gather_output)
        if getattr(self.config, "use_mup", False) and self.config.mup_output_mult != 1.0:
            logits = logits * self.config.mup_output_mult            # muP: logits stay O(1) as the width grows
        if labels is None:
            return logits.transpose(0, 1).contiguous()  # [b, s, v/tp]
        return self.compute_language_model_loss(labels, logits)

    def shared_embedding_or_output_weight(self) -> Tensor:
        if self.pre_process:
            return self.embedding.word_embeddings.weight
        if self.post_process:
            return self.output_layer.weight
        return None

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: Optional[Dict] = None) -> ShardedStateDict:
        return super().sharded_state_dict(prefix, sharded_offsets, metadata)
