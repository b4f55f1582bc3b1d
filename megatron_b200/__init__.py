"""megatron_b200 — a Blackwell-native (sm_100a / NVLink 5) Megatron-Core-equivalent framework."""
from .core.package_info import __version__
