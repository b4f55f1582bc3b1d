from ...ssm.mamba_hybrid_layer_allocation import Symbols, allocate_layers, parse_hybrid_pattern  # noqa: F401
