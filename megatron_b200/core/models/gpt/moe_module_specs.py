"""MoE layer specs (reference ``models/gpt/moe_module_specs.py``)."""
from typing import Optional

from ...transformer.mlp import MLPSubmodules
from ...transformer.moe.moe_layer import MoELayer, MoESubmodules
from ...transformer.moe.shared_experts import SharedExpertMLP
from ...transformer.spec_utils import ModuleSpec


def get_moe_module_spec_for_backend(backend, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False, moe_use_legacy_grouped_gemm: bool = False):
    assert num_experts is not None
    experts_cls, subs = backend.grouped_mlp_modules(moe_grouped_gemm, moe_use_legacy_grouped_gemm)
    experts = ModuleSpec(module=experts_cls, submodules=subs)
    shared = ModuleSpec(module=SharedExpertMLP, params={"gate": False},
                        submodules=MLPSubmodules(linear_fc1=backend.column_parallel_linear(), linear_fc2=backend.row_parallel_linear()))
    return ModuleSpec(module=MoELayer, submodules=MoESubmodules(experts=experts, shared_experts=shared))


def get_moe_module_spec(use_te: bool = False, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False, moe_use_legacy_grouped_gemm: bool = False):
    from ..backends import B200SpecProvider

    return get_moe_module_spec_for_backend(B200SpecProvider(), num_experts, moe_grouped_gemm, moe_use_legacy_grouped_gemm)
