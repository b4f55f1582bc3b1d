"""Base class for language models: loss, embedding tying across pipeline stages
(reference ``models/common/language_module/language_module.py:161,208-324``)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from megatron_b200.core import parallel_state as ps
from megatron_b200.core.tensor_parallel import vocab_parallel_cross_entropy
from megatron_b200.core.transformer.module import MegatronModule
from megatron_b200.core.transformer.transformer_config import TransformerConfig
from megatron_b200.core.utils import make_tp_sharded_tensor_for_checkpoint


class LanguageModule(MegatronModule):
    def __init__(self, config: TransformerConfig, pg_collection=None):
        super().__init__(config)
        self.pg_collection = pg_collection
        self.vp_stage = None

    def compute_language_model_loss(self, labels: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        """labels [b, s]; logits [s, b, v/tp] → loss [b, s]."""
        labels = labels.transpose(0, 1).contiguous()
        tp = getattr(self.pg_collection, "tp", None) if self.pg_collection is not None else None
        loss = vocab_parallel_cross_entropy(logits, labels, tp_group=tp)
        return loss.transpose(0, 1).contiguous()

    def setup_embeddings_and_output_layer(self) -> None:
        """Tied embeddings with pp>1: the last stage holds a zero-initialised copy of the
        word embeddings; ``finalize_model_grads`` all-reduces the two grads over the
        embedding group so both copies stay identical."""
        if self.pre_process:
            self.embedding.word_embeddings.weight.is_embedding_or_output_parameter = True
        if self.post_process and getattr(self, "output_layer", None) is not None and self.output_layer.weight is not None:
            self.output_layer.weight.is_embedding_or_output_parameter = True
        if not self.share_embeddings_and_output_weights:
            return
        if ps.get_pipeline_model_parallel_world_size() == 1:
            self.shared_embedding_or_output_weight().zero_out_wgrad = True
            return
        if ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=self.vp_stage) and self.pre_process and not self.post_process:
            self.shared_embedding_or_output_weight().shared_embedding = True
        if self.post_process and not self.pre_process:
            assert not ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=self.vp_stage)
            self.output_layer.weight.data.fill_(0)
            self.output_layer.weight.shared = True
            self.output_layer.weight.shared_embedding = True
        if torch.distributed.is_initialized() and ps.is_rank_in_embedding_group(ignore_virtual=False, vp_stage=self.vp_stage):
            w = self.shared_embedding_or_output_weight()
            if w.is_cuda or torch.distributed.get_backend() == "gloo":
                torch.distributed.all_reduce(w.data, group=ps.get_embedding_group())

    def shared_embedding_or_output_weight(self) -> Optional[torch.Tensor]:
        if self.pre_process:
            return self.embedding.word_embeddings.weight
        if self.post_process:
            return self.output_layer.weight
        return None

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: Tuple = (), metadata: Optional[dict] = None):
        assert not sharded_offsets, "unexpected sharded offsets"
        sd = super().sharded_state_dict(prefix, sharded_offsets, metadata)
        first_key = f"{prefix}embedding.word_embeddings.weight"
        out_w = f"{prefix}output_layer.weight"
        out_extra = f"{prefix}output_layer._extra_state"
        if self.share_embeddings_and_output_weights:
            self._tie_in_sharded_state_dict(sd, out_w, first_key, metadata)
        elif self.post_process and out_w in sd:
            sd[out_w].allow_shape_mismatch = True
        # GPT checkpoints never store an extra state for the output layer (reference gpt_model.py sharded_state_dict pops it after checking it is empty)
        extra = sd.pop(out_extra, None)
        assert not (extra is not None and getattr(extra, "data", None)), f"expected the output layer extra state to be empty, got {extra}"
        return sd

    def _tie_in_sharded_state_dict(self, sd, output_layer_weight_key, first_stage_word_emb_key, metadata=None):
        if not self.post_process:
            sd.pop(output_layer_weight_key, None)
            return
        if self.pre_process:
            sd.pop(output_layer_weight_key, None)
            return
        # last stage, tied: save under the embedding key as a *replica* of the first-stage copy
        t = self.shared_embedding_or_output_weight()
        dp_rank = ps.get_data_parallel_rank(with_context_parallel=True)
        sd[output_layer_weight_key] = make_tp_sharded_tensor_for_checkpoint(
            t, first_stage_word_emb_key, replica_id=(1, 0, dp_rank), allow_shape_mismatch=True
        )
