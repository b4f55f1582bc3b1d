from .hybrid_optimizer import HybridDeviceOptimizer  # noqa: F401
