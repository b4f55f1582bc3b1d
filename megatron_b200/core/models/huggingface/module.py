"""Hugging Face models as Megatron modules (reference ``models/huggingface/{module,clip_model,qwen_model}.py``): lets a multimodal model use an HF
vision tower or language model as a sub-module with the Megatron module contract (``set_input_tensor``, ``config``, plain ``state_dict``).
No network is assumed: ``build_hf_model`` takes a local path or an in-memory ``transformers`` config."""
from __future__ import annotations

from typing import Optional

import torch

from ...transformer.module import MegatronModule


class HuggingFaceModule(MegatronModule):
    """Base wrapper: single-stage (no PP) module whose parameters are ordinary replicated tensors."""

    def __init__(self, config, hf_model: Optional[torch.nn.Module] = None):
        super().__init__(config=config)
        self.model = hf_model
        self.input_tensor = None

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == "model" and isinstance(value, torch.nn.Module):
            for p in value.parameters():        # replicated across TP: reduce like layer norms under sequence parallelism
                setattr(p, "sequence_parallel", bool(getattr(self.config, "sequence_parallel", False)))


class AutoHuggingFaceModel(HuggingFaceModule):
    """``forward(*args, **kwargs)`` → the HF model's ``last_hidden_state`` (or logits / first output when it has none)."""

    def __init__(self, config, hf_model: torch.nn.Module, output_key: Optional[str] = None):
        super().__init__(config, hf_model)
        self.output_key = output_key

    def forward(self, *args, **kwargs):
        out = self.model(*args, **kwargs)
        if self.output_key is not None:
            return out[self.output_key] if isinstance(out, dict) or hasattr(out, "keys") else getattr(out, self.output_key)
        for k in ("last_hidden_state", "logits"):
            if hasattr(out, k) and getattr(out, k) is not None:
                return getattr(out, k)
        return out[0] if isinstance(out, (tuple, list)) else out


def build_hf_model(config, model_name_or_path: Optional[str] = None, hf_config=None, model_cls: str = "AutoModel", **kwargs) -> AutoHuggingFaceModel:
    """From a local checkpoint directory (``from_pretrained``, offline) or from a ``transformers`` config object (random init — tests, pre-training)."""
    import transformers

    cls = getattr(transformers, model_cls)
    if hf_config is not None:
        hf = cls.from_config(hf_config) if hasattr(cls, "from_config") else cls(hf_config)
    else:
        hf = cls.from_pretrained(model_name_or_path, local_files_only=True, **kwargs)
    if getattr(config, "bf16", False):
        hf = hf.to(torch.bfloat16)
    return AutoHuggingFaceModel(config, hf)
