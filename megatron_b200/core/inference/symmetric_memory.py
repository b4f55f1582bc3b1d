"""Named symmetric-memory buffers for inference collectives (reference ``inference/symmetric_memory.py`` — ``SymmetricMemoryManager`` :133).

Latency-bound decode collectives (a [b, h] all-reduce per layer) want pre-registered peer-mapped buffers instead of a registration per call.
The manager hands out named, shape-keyed tensors carved from the NVLink backend's symmetric heap (``parallel/nvlink.py``: VMM allocation, peer
mapping, NVLS multicast alias) and falls back to ordinary device tensors + NCCL when symmetric memory is unavailable (CPU tests, no NVSwitch)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


class SymmetricMemoryBuffer:
    def __init__(self, tensor: torch.Tensor, backend=None, offset: int = 0):
        self.tensor, self.backend, self.offset = tensor, backend, offset

    @property
    def is_symmetric(self) -> bool:
        return self.backend is not None

    def all_reduce_(self, group=None) -> torch.Tensor:
        if self.backend is not None and hasattr(self.backend, "all_reduce"):
            return self.backend.all_reduce(self.tensor)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.tensor, group=group)
        return self.tensor


class SymmetricMemoryManager:
    _instances: Dict[int, "SymmetricMemoryManager"] = {}

    def __init__(self, group=None, max_bytes: int = 64 << 20):
        self.group, self.max_bytes = group, max_bytes
        self.buffers: Dict[Tuple, SymmetricMemoryBuffer] = {}
        self.backend = None
        self.used = 0
        if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl" and dist.get_world_size(group) > 1:
            try:
                from ...parallel import collectives

                self.backend = collectives.enable_for_group(group if group is not None else dist.group.WORLD)
            except Exception:
                self.backend = None

    @classmethod
    def get(cls, group=None) -> "SymmetricMemoryManager":
        key = id(group)
        if key not in cls._instances:
            cls._instances[key] = cls(group)
        return cls._instances[key]

    def get_buffer(self, name: str, shape, dtype: torch.dtype = torch.bfloat16, device=None) -> SymmetricMemoryBuffer:
        key = (name, tuple(shape), dtype)
        buf = self.buffers.get(key)
        if buf is None:
            nbytes = int(torch.tensor(shape).prod()) * torch.empty(0, dtype=dtype).element_size()
            if self.used + nbytes > self.max_bytes:
                raise MemoryError(f"symmetric buffer '{name}' ({nbytes} B) exceeds the {self.max_bytes} B budget ({self.used} B in use)")
            t = None
            if self.backend is not None and hasattr(self.backend, "alloc_symmetric"):
                t = self.backend.alloc_symmetric(shape, dtype)
            if t is None:
                dev = device or ("cuda" if torch.cuda.is_available() else "cpu")
                t = torch.zeros(tuple(shape), dtype=dtype, device=dev)
                buf = SymmetricMemoryBuffer(t, None)
            else:
                buf = SymmetricMemoryBuffer(t, self.backend)
            self.used += nbytes
            self.buffers[key] = buf
        return buf

    def release_all(self) -> None:
        self.buffers.clear()
        self.used = 0
