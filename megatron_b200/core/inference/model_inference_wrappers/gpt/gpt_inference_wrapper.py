"""Decoder-only wrapper (reference ``model_inference_wrappers/gpt/gpt_inference_wrapper.py``)."""
from typing import Any, Dict

import torch

from ..abstract_model_inference_wrapper import AbstractModelInferenceWrapper


class GPTInferenceWrapper(AbstractModelInferenceWrapper):
    def prep_inference_input(self, prompts_tokens: torch.Tensor) -> Dict[str, Any]:
        """``prompts_tokens`` [b, max_len] (padded prompts + room for the generated tokens) → tokens and position ids; attention is causal in the fused path, so no
        dense mask is materialised (``attention_mask`` stays None)."""
        b, s = prompts_tokens.shape
        pos = torch.arange(s, dtype=torch.long, device=prompts_tokens.device).unsqueeze(0).expand(b, -1)
        return {"tokens": prompts_tokens, "attention_mask": None, "position_ids": pos}

    def get_batch_for_context_window(self, inference_input: Dict[str, Any], context_start_position: int, context_end_position: int) -> Dict[str, Any]:
        sl = slice(context_start_position, context_end_position)
        return {"tokens": inference_input["tokens"][:, sl], "position_ids": inference_input["position_ids"][:, sl], "attention_mask": None}
