from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import MLATransformerConfig, TransformerConfig
from .transformer_layer import TransformerLayer, TransformerLayerSubmodules

__all__ = ["MegatronModule", "ModuleSpec", "build_module", "TransformerConfig", "MLATransformerConfig", "TransformerLayer", "TransformerLayerSubmodules"]
from .transformer_layer import HyperConnectionTransformerLayer  # noqa: E402,F401
