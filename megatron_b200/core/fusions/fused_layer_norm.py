"""LayerNorm / RMSNorm entry point (reference ``fusions/fused_layer_norm.py:30-167``) → ``FusedNorm`` (``csrc/norm.cu``)."""
from ..transformer.torch_norm import FusedNorm as FusedLayerNorm  # noqa: F401
