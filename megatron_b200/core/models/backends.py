"""Spec providers: which classes implement each building block.

``B200SpecProvider`` is the default (fused NVLink linears, sm_100a norm/attention
kernels).  ``LocalSpecProvider`` keeps the reference's name for the same set —
there is no TransformerEngine provider (reference ``models/backends.py:99-147``).
"""
from __future__ import annotations

from typing import Optional, Tuple

from ..tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
from ..transformer.dot_product_attention import DotProductAttention
from ..transformer.torch_norm import FusedNorm


class BackendSpecProvider:
    def linear(self):
        raise NotImplementedError

    def column_parallel_linear(self):
        raise NotImplementedError

    def row_parallel_linear(self):
        raise NotImplementedError

    def fuse_layernorm_and_linear(self) -> bool:
        return False

    def column_parallel_layer_norm_linear(self):
        return None

    def layer_norm(self, rms_norm: bool = False, for_qk: bool = False):
        raise NotImplementedError

    def core_attention(self):
        raise NotImplementedError

    def grouped_mlp_modules(self, moe_use_grouped_gemm: bool, moe_use_legacy_grouped_gemm: bool = False):
        raise NotImplementedError

    def activation_func(self):
        return None


class B200SpecProvider(BackendSpecProvider):
    def linear(self):
        return ColumnParallelLinear

    def column_parallel_linear(self):
        return ColumnParallelLinear

    def row_parallel_linear(self):
        return RowParallelLinear

    def layer_norm(self, rms_norm: bool = False, for_qk: bool = False):
        return FusedNorm

    def core_attention(self):
        return DotProductAttention

    def grouped_mlp_modules(self, moe_use_grouped_gemm: bool, moe_use_legacy_grouped_gemm: bool = False) -> Tuple[type, Optional[object]]:
        from ..transformer.mlp import MLPSubmodules
        from ..transformer.moe.experts import GroupedMLP, SequentialMLP

        subs = MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear)
        return (GroupedMLP if moe_use_grouped_gemm else SequentialMLP), subs


LocalSpecProvider = B200SpecProvider
