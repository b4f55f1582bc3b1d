from .heterogeneous_config import HeterogeneousTransformerConfig, get_gpt_heterogeneous_layer_spec  # noqa: F401
