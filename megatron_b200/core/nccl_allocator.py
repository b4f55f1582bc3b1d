"""Communication-registered memory (reference ``core/nccl_allocator.py:31-381``, N3).

The reference plugs ``ncclMemAlloc`` into a ``torch.cuda.MemPool`` and registers the pool with each process group so NCCL can use
NVLS / zero-copy on those buffers.  This framework's collectives run over a *symmetric heap* (``parallel/nvlink.py``): one VMM
allocation per rank, peer-mapped on every rank of the group, with an NVLS multicast alias.  The API below keeps the reference's
shape (``init`` / ``nccl_mem`` / ``MultiGroupMemPoolAllocator``) and routes allocations to that heap, so DDP grad/param buffers and
TP workspaces created inside ``with nccl_mem(group):`` are directly usable by the multimem kernels."""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch

_BACKENDS: Dict[int, object] = {}


def init() -> None:
    """Kept for API parity (the reference sets ``TORCH_NCCL_USE_TENSOR_REGISTER_ALLOCATOR_HOOK``); nothing to configure here."""


def _backend(group):
    from ..parallel import collectives

    key = id(group)
    if key not in _BACKENDS:
        _BACKENDS[key] = collectives.enable_for_group(group)
    return _BACKENDS[key]


class SymmetricAllocation:
    """Handle returned by the ``nccl_mem`` context: ``alloc(numel, dtype)`` carves tensors out of the group's symmetric heap."""

    def __init__(self, groups: List):
        self.groups = groups
        self.tensors: List[torch.Tensor] = []

    def alloc(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        if not torch.cuda.is_available() or not self.groups:
            t = torch.zeros(numel, dtype=dtype)
        else:
            t = _backend(self.groups[0]).alloc_symmetric(numel, dtype)
        self.tensors.append(t)
        return t


@contextlib.contextmanager
def nccl_mem(pool=None, enabled: bool = True, device=None, group=None, symmetric: bool = True):
    """``with nccl_mem(group=g) as mem: buf = mem.alloc(n, torch.float32)``.  With ``enabled=False`` plain device memory is used."""
    yield SymmetricAllocation([group] if (enabled and group is not None) else [])


class MultiGroupMemPoolAllocator:
    """Allocate buffers that are registered with SEVERAL groups (reference ``:276``).  A symmetric allocation belongs to exactly one
    rendezvous; for multiple groups the buffer is rendezvoused on the largest group, whose peer mappings cover the sub-groups."""

    def __init__(self, pool=None, groups: Optional[List] = None, symmetric: bool = True):
        self.groups = sorted(groups or [], key=lambda g: -torch.distributed.get_world_size(g)) if groups else []
        self._alloc = SymmetricAllocation(self.groups[:1])

    def __enter__(self):
        return self._alloc

    def __exit__(self, *exc):
        return False


def create_nccl_mem_pool(symmetric: bool = True):
    """Parity shim: the symmetric heap is created lazily per group; returns None."""
    return None
