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


    # ---- batched decode: every running request advances by one token in ONE forward --------------------------------
    def batch_tables(self, rids: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """→ (block table ``[B, max_blocks]`` padded with block 0, current lengths ``[B]``) on the cache's device."""
        dev = self.k.device
        width = max(len(self.block_tables[r]) for r in rids)
        table = torch.zeros(len(rids), width, dtype=torch.long)
        for i, r in enumerate(rids):
            t = self.block_tables[r]
            table[i, : len(t)] = torch.tensor(t)
        return table.to(dev), torch.tensor([self.lengths[r] for r in rids], dtype=torch.long, device=dev)

    def append_batch(self, layer: int, table: torch.Tensor, positions: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> None:
        """``k, v [B, kv_heads, head_dim]``: request ``i``'s new entry goes to its position ``positions[i]``."""
        blk = table.gather(1, (positions // self.block_size).unsqueeze(1)).squeeze(1)
        off = positions % self.block_size
        self.k[layer, blk, off] = k
        self.v[layer, blk, off] = v

    def gather_batch(self, layer: int, table: torch.Tensor, max_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """→ ``K, V [B, L, kv_heads, head_dim]`` with ``L = ceil(max_len / block) · block``; rows past a request's length are whatever the
        padded blocks hold — the caller masks them."""
        nblk = (max_len + self.block_size - 1) // self.block_size
        t = table[:, :nblk]
        B = t.shape[0]
        return self.k[layer][t].reshape(B, nblk * self.block_size, *self.k.shape[3:]), self.v[layer][t].reshape(B, nblk * self.block_size, *self.v.shape[3:])


class BatchedDecodeContext:
    """Inference context of one batched decode step (understood by ``Attention.forward``): request ``i`` of the batch sits at
    position ``lengths[i]``, its K/V history lives in the paged cache behind row ``i`` of ``block_table``."""

    is_batched_decode = True

    def __init__(self, cache: PagedKVCache, rids: List[int], layer_numbers: List[int]):
        self.cache, self.rids = cache, rids
        self.block_table, self.lengths = cache.batch_tables(rids)
        self.max_len = int(max(cache.lengths[r] for r in rids)) + 1
        self.max_sequence_length = self.max_len                      # rotary table length
        self.max_batch_size = len(rids)
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict: Dict = {}
        self.layer_index = {n: i for i, n in enumerate(layer_numbers)}
