"""Deprecated holder of wrapper options (reference ``model_inference_wrappers/inference_wrapper_config.py``); the context and the model config carry them now."""
from dataclasses import dataclass

import torch


@dataclass
class InferenceWrapperConfig:
    hidden_size: int
    params_dtype: torch.dtype
    inference_batch_times_seqlen_threshold: int
    padded_vocab_size: int
    inference_max_requests: int = 8
    inference_max_seq_length: int = 2560
    fp32_residual_connection: bool = False
    nccl_all_reduce_for_prefill: bool = False

    def add_attributes(self, attribute_value_pair: dict) -> None:
        for k, v in attribute_value_pair.items():
            setattr(self, k, v)
