"""Precision policy of Megatron-FSDP (reference ``megatron_fsdp/mixed_precision.py:1-407``).

The reference's fp8 helpers wrap TransformerEngine's ``Float8Tensor``; here a quantised parameter is the pair this framework's
block-scaled GEMMs consume — an E4M3 payload (uint8) plus E8M0 scales per 32 elements (``ops.extra.mxfp8_quantize``) — carried by
``QuantizedParam``.  FSDP shards and all-gathers the RAW payload bytes (half the bf16 traffic) and gathers scales alongside."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch


@dataclass
class MixedPrecisionPolicy:
    """``main_params_dtype``: optimizer copy (None = same as the model weights, no extra copy); ``main_grads_dtype``: accumulation
    and reduction dtype of the gradient shards; ``grad_comm_dtype``: wire dtype of the reduce-scatter (None = ``main_grads_dtype``)."""
    main_params_dtype: Optional[torch.dtype] = torch.float32
    main_grads_dtype: Optional[torch.dtype] = torch.float32
    grad_comm_dtype: Optional[torch.dtype] = None
    fp8_param_gather: bool = False

    def preserve_fp32_weights(self) -> bool:
        return self.main_params_dtype == torch.float32

    def grad_reduce_in_fp32(self) -> bool:
        return (self.grad_comm_dtype or self.main_grads_dtype) == torch.float32


class QuantizedParam:
    """MXFP8 weight: ``payload`` uint8 [rows, K], ``scales`` uint8 [rows, K/32]; optional cached transpose for the wgrad GEMM."""

    def __init__(self, payload: torch.Tensor, scales: torch.Tensor, shape: Tuple[int, ...]):
        self.payload, self.scales, self.shape = payload, scales, tuple(shape)
        self.transpose_cache: Optional[Tuple[torch.Tensor, torch.Tensor]] = None


def is_float8tensor(t) -> bool:
    return isinstance(t, QuantizedParam)


def is_blockwise_float8tensor(t) -> bool:
    return isinstance(t, QuantizedParam)                     # every quantised parameter here is block-scaled (1x32)


def fp8_quantize(x: torch.Tensor) -> QuantizedParam:
    from ..... import ops
    q, sf = ops.extra.mxfp8_quantize(x.reshape(-1, x.shape[-1]))
    return QuantizedParam(q, sf, x.shape)


def fp8_dequantize(t: QuantizedParam) -> torch.Tensor:
    from ..... import ops
    return ops.extra.mxfp8_dequantize(t.payload, t.scales).view(t.shape)


def fp8_get_raw_data(t: QuantizedParam, transpose: bool = False) -> torch.Tensor:
    if transpose:
        assert t.transpose_cache is not None, "no transpose cache: call fp8_create_transpose_cache first"
        return t.transpose_cache[0]
    return t.payload


def fp8_set_raw_data(t: QuantizedParam, data: torch.Tensor, set_transpose: bool = False) -> None:
    if set_transpose:
        t.transpose_cache = (data, t.transpose_cache[1] if t.transpose_cache else None)
    else:
        assert data.numel() == t.payload.numel()
        t.payload = data.view(t.payload.shape)


def fp8_need_transpose_data(t: QuantizedParam) -> bool:
    """1x32 scales along K are not transposable: the wgrad GEMM needs the weight quantised along the OTHER axis."""
    return True


def fp8_need_transpose_data_for_meta_device_init(module) -> bool:
    return True


def fp8_create_transpose_cache(t: QuantizedParam) -> None:
    from ..... import ops
    w = fp8_dequantize(t)
    wt = w.reshape(-1, w.shape[-1]).t().contiguous()
    t.transpose_cache = ops.extra.mxfp8_quantize(wt)


def fp8_discard_transpose_cache(t: QuantizedParam) -> None:
    t.transpose_cache = None


def get_quantized_model_init_context_cls():
    """Context under which linears create quantised weights directly (no bf16 master on the device)."""
    from contextlib import nullcontext
    return nullcontext
