from ..kv_cache import KVBlockAllocator  # noqa: F401
