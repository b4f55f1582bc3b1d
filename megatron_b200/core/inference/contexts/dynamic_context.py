"""Continuous-batching context (reference ``inference/contexts/dynamic_context.py``): in this framework the paged KV cache owns the block tables and the
per-step views the attention layers read (``kv_cache.PagedKVCache`` → ``PagedPrefillContext`` / ``BatchedDecodeContext``)."""
from ..kv_cache import BatchedDecodeContext, KVBlockAllocator, PagedKVCache, PagedPrefillContext  # noqa: F401
from .base_context import BaseInferenceContext


class DynamicInferenceContext(PagedKVCache, BaseInferenceContext):
    """The paged cache under the reference's name.  ``prefill_view`` / ``decode_view`` hand the attention layers their per-step contexts."""

    def is_static_batching(self) -> bool:
        return False


class ContextOverflowError(RuntimeError):
    """A request does not fit: no free KV blocks / over the token or request budget."""


class RequestOverflowError(ContextOverflowError):
    pass


class TokenOverflowError(ContextOverflowError):
    pass


class MaxSequenceLengthOverflowError(ContextOverflowError):
    pass


class BlockOverflowError(ContextOverflowError):
    pass
