"""Managed (unified) memory pool for inference state (reference ``inference/unified_memory.py:87-240``, N5).

The reference compiles an inline C++ string with ``load_inline``; here the two allocator entry points
(``mb200_managed_malloc`` / ``mb200_managed_free``, ``ops/csrc/runtime_native.cu``) live in the prebuilt in-tree extension and are
handed to ``torch.cuda.memory.CUDAPluggableAllocator``.  KV-cache tensors allocated inside ``unified_memory_pool()`` may exceed free
HBM: pages prefer the GPU and migrate on demand."""
from __future__ import annotations

import contextlib
import warnings
from enum import Enum, auto

import torch


class CompilationState(Enum):
    UNATTEMPTED = auto()
    FAILURE = auto()
    SUCCESS = auto()


_state = CompilationState.UNATTEMPTED
_alloc = None
_pool = None


class UnifiedMemoryUnsupportedError(RuntimeError):
    pass


def _load():
    global _state, _alloc
    if _state != CompilationState.UNATTEMPTED:
        return
    try:
        from ...ops.build import target_path

        so = str(target_path())
        _alloc = torch.cuda.memory.CUDAPluggableAllocator(so, "mb200_managed_malloc", "mb200_managed_free")
        _state = CompilationState.SUCCESS
    except Exception as e:  # pragma: no cover - needs a GPU build
        warnings.warn(f"unified-memory allocator unavailable: {e}")
        _state = CompilationState.FAILURE


def has_unified_memory() -> bool:
    _load()
    return _state == CompilationState.SUCCESS


def create_unified_mempool():
    """``torch.cuda.MemPool`` whose blocks come from ``cudaMallocManaged``."""
    global _pool
    _load()
    if _state != CompilationState.SUCCESS:
        raise UnifiedMemoryUnsupportedError("managed-memory allocator could not be loaded")
    if _pool is None:
        _pool = torch.cuda.MemPool(_alloc.allocator())
    return _pool


@contextlib.contextmanager
def unified_memory_pool():
    """Allocate tensors created inside the block from managed memory."""
    pool = create_unified_mempool()
    with torch.cuda.use_mem_pool(pool):
        yield pool
