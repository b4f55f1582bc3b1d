from ..mamba.mamba_model import MambaModel as HybridModel  # noqa: F401
