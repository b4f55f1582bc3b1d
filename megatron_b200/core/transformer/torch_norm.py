"""Normalisation layers that work under sequence parallelism.

The reference's local backend refuses RMSNorm + sequence-parallel
(``transformer/torch_norm.py:52``); here the norm weight is tagged
``sequence_parallel`` so ``finalize_model_grads`` all-reduces its gradient over
the TP group, and the math runs in the sm_100a RMSNorm/LayerNorm kernels.
"""
from __future__ import annotations

import torch

from ... import ops
from .transformer_config import TransformerConfig


class FusedNorm(torch.nn.Module):
    """RMSNorm / LayerNorm with optional zero-centred gamma (``weight`` stores ``gamma - 1``)."""

    def __init__(self, config: TransformerConfig, hidden_size: int, eps: float = 1e-5, persist_layer_norm: bool = False,
                 zero_centered_gamma: bool = False, normalization: str = None, **kwargs):
        super().__init__()
        self.config = config
        self.normalization = normalization or config.normalization
        assert self.normalization in ("LayerNorm", "RMSNorm"), self.normalization
        self.eps = eps
        self.zero_centered_gamma = config.layernorm_zero_centered_gamma or zero_centered_gamma
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        init = 0.0 if self.zero_centered_gamma else 1.0
        self.weight = torch.nn.Parameter(torch.full((hidden_size,), init, dtype=config.params_dtype, device=dev))
        if self.normalization == "LayerNorm":
            self.bias = torch.nn.Parameter(torch.zeros(hidden_size, dtype=config.params_dtype, device=dev))
        else:
            self.register_parameter("bias", None)
        sp = config.sequence_parallel
        setattr(self.weight, "sequence_parallel", sp)
        if self.bias is not None:
            setattr(self.bias, "sequence_parallel", sp)

    def forward(self, x):
        if self.normalization == "RMSNorm":
            return ops.rms_norm(x, self.weight, self.eps, self.zero_centered_gamma)
        return ops.layer_norm(x, self.weight, self.bias, self.eps, self.zero_centered_gamma)


class WrappedTorchNorm:
    """Factory kept for spec compatibility (reference ``WrappedTorchNorm``)."""

    def __new__(cls, config, hidden_size, eps=1e-5, persist_layer_norm=False, zero_centered_gamma=False, normalization="LayerNorm", **kw):
        return FusedNorm(config, hidden_size, eps, persist_layer_norm, zero_centered_gamma, normalization or config.normalization)


class L2Norm(torch.nn.Module):
    """Parameter-free l2 normalisation over the last dim (QK-l2norm)."""

    def __init__(self, config=None, hidden_size=None, eps: float = 1e-6, **kwargs):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)


LayerNormImpl = FusedNorm
