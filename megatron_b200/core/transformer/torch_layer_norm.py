"""``WrappedTorchLayerNorm`` under the reference's module path (``transformer/torch_layer_norm.py`` is itself a deprecated alias of ``torch_norm``)."""
from .torch_norm import *  # noqa: F401,F403
from .torch_norm import WrappedTorchNorm as WrappedTorchLayerNorm  # noqa: F401
