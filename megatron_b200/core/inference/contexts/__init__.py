from .base_context import BaseInferenceContext  # noqa: F401
from .dynamic_context import (BlockOverflowError, ContextOverflowError, DynamicInferenceContext, MaxSequenceLengthOverflowError, RequestOverflowError,  # noqa: F401
                              TokenOverflowError)
from .kv_block_allocator import KVBlockAllocator  # noqa: F401
from .static_context import StaticInferenceContext  # noqa: F401
