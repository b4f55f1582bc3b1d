"""Hybrid (Mamba / gated-delta-net / attention / MLP / MoE) language model under the reference's newer package name (``models/hybrid``); ``models/mamba`` keeps
the original names."""
from .hybrid_block import HybridStack, HybridStackSubmodules  # noqa: F401
from .hybrid_layer_allocation import Symbols, allocate_layers, parse_hybrid_pattern  # noqa: F401
from .hybrid_layer_specs import hybrid_stack_spec  # noqa: F401
from .hybrid_model import HybridModel  # noqa: F401
