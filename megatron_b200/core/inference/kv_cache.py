"""Paged KV cache with a block allocator (reference ``inference/contexts/dynamic_context.py:299``,
``KVBlockAllocator``): fixed-size blocks in one pool per layer, a block table per request, so
requests of different lengths share memory without fragmentation and can be added/evicted at any step."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


class KVBlockAllocator:
    def __init__(self, num_blocks: int):
        self.num_blocks = num_blocks
        self.free: List[int] = list(range(num_blocks - 1, -1, -1))

    def allocate(self, n: int) -> Optional[List[int]]:
        if n > len(self.free):
            return None
        return [self.free.pop() for _ in range(n)]

    def release(self, blocks: List[int]):
        self.free.extend(reversed(blocks))

    @property
    def num_free(self) -> int:
        return len(self.free)


class PagedKVCache:
    """``k/v`` pools: [layers, num_blocks, block_size, kv_heads, head_dim]."""

    def __init__(self, num_layers: int, num_blocks: int, block_size: int, kv_heads: int, head_dim: int, dtype, device):
        self.block_size = block_size
        self.k = torch.zeros(num_layers, num_blocks, block_size, kv_heads, head_dim, dtype=dtype, device=device)
        self.v = torch.zeros_like(self.k)
        self.allocator = KVBlockAllocator(num_blocks)
        self.block_tables: Dict[int, List[int]] = {}
        self.lengths: Dict[int, int] = {}

    def can_admit(self, num_tokens: int) -> bool:
        return self.allocator.num_free >= (num_tokens + self.block_size - 1) // self.block_size

    def add_request(self, rid: int, num_tokens: int) -> bool:
        blocks = self.allocator.allocate((num_tokens + self.block_size - 1) // self.block_size)
        if blocks is None:
            return False
        self.block_tables[rid], self.lengths[rid] = blocks, 0
        return True

    def ensure_capacity(self, rid: int, new_len: int) -> bool:
        need = (new_len + self.block_size - 1) // self.block_size - len(self.block_tables[rid])
        if need > 0:
            blocks = self.allocator.allocate(need)
            if blocks is None:
                return False
            self.block_tables[rid].extend(blocks)
        return True

    def release(self, rid: int):
        self.allocator.release(self.block_tables.pop(rid))
        self.lengths.pop(rid)

    def append(self, layer: int, rid: int, k: torch.Tensor, v: torch.Tensor, start: int):
        """k, v: [n, kv_heads, head_dim] for positions [start, start+n) of request ``rid``."""
        n = k.shape[0]
        pos = torch.arange(start, start + n, device=k.device)
        table = torch.tensor(self.block_tables[rid], device=k.device)
        blk, off = table[pos // self.block_size], pos % self.block_size
        self.k[layer, blk, off] = k
        self.v[layer, blk, off] = v

    def gather(self, layer: int, rid: int, length: int) -> Tuple[torch.Tensor, torch.Tensor]:
        table = torch.tensor(self.block_tables[rid], device=self.k.device)
        nblk = (length + self.block_size - 1) // self.block_size
        k = self.k[layer, table[:nblk]].reshape(-1, *self.k.shape[3:])[:length]
        v = self.v[layer, table[:nblk]].reshape(-1, *self.v.shape[3:])[:length]
        return k, v
