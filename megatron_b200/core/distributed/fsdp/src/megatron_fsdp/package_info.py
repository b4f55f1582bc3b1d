"""Package metadata (reference ``megatron_fsdp/package_info.py``)."""
MAJOR, MINOR, PATCH = 0, 2, 0
__version__ = f"{MAJOR}.{MINOR}.{PATCH}"
__package_name__ = "megatron_fsdp"
__description__ = "Megatron-FSDP for B200: bucketed ZeRO-1/2/3 with double-buffered fixed pools and per-unit prefetch"
