"""Model-side half of the static generation loop (reference ``model_inference_wrappers/abstract_model_inference_wrapper.py``): prepare the inputs of a batch of
prompts once, slice them per context window, run one forward step against the KV cache in the inference context and hand back the logits."""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional

import torch

from ... import parallel_state as ps
from ..contexts import BaseInferenceContext, StaticInferenceContext


class AbstractModelInferenceWrapper(abc.ABC):
    def __init__(self, model, inference_context: Optional[BaseInferenceContext] = None, pg_collection=None):
        assert not isinstance(model, (list, tuple)), "interleaved pipeline schedules are a training feature: pass one model"
        self.model = model
        cfg = getattr(model, "config", None)
        self.inference_context = inference_context or StaticInferenceContext(getattr(cfg, "inference_max_requests", None) or 8, getattr(cfg, "inference_max_seq_length", None) or
                                                                              getattr(model, "max_sequence_length", 2048))
        self.pg_collection = pg_collection
        if ps.model_parallel_is_initialized() and ps.get_pipeline_model_parallel_world_size() > 1:
            raise NotImplementedError("the wrapper drives one pipeline stage; pipelined serving goes through the engines")

    def prep_model_for_inference(self) -> None:
        self.model.eval()
        self.inference_context.reset()

    @abc.abstractmethod
    def prep_inference_input(self, prompts_tokens: torch.Tensor) -> Dict[str, Any]:
        ...

    @abc.abstractmethod
    def get_batch_for_context_window(self, inference_input: Dict[str, Any], context_start_position: int, context_end_position: int) -> Dict[str, Any]:
        ...

    def _forward(self, inference_input: Dict[str, Any]) -> torch.Tensor:
        return self.model(inference_input["tokens"], inference_input["position_ids"], inference_input.get("attention_mask"), inference_context=self.inference_context,
                          runtime_gather_output=True)

    @torch.no_grad()
    def run_one_forward_step(self, inference_input: Dict[str, Any], recv_buffer_seq_len: Optional[int] = None) -> torch.Tensor:
        """Logits ``[b, s, vocab]`` of this window; the sequence offset of the context advances by the window length."""
        tokens = inference_input["tokens"]
        self.inference_context.current_batch_size = tokens.shape[0]
        logits = self._forward(inference_input)
        self.inference_context.increment_sequence_len_offset(tokens.shape[1])
        return logits
