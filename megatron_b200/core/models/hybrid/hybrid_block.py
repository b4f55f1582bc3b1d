from ...ssm.mamba_block import MambaStack as HybridStack, MambaStackSubmodules as HybridStackSubmodules  # noqa: F401
