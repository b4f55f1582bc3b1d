"""Paged KV cache with a block allocator (reference ``inference/contexts/dynamic_context.py:299``,
``KVBlockAllocator``): fixed-size blocks in one pool per layer, a block table per request, so
requests of different lengths share memory without fragmentation and can be added/evicted at any step."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


class KVBlockAllocator:
    """Free list + (optional) prefix cache (reference ``contexts/kv_block_allocator.py``: ``enable_prefix_caching``, LRU / ref-zero eviction).

    With prefix caching a FULL block whose content is determined by the token prefix that produced it is registered under the chained
    hash of that prefix.  Blocks are reference counted; when the last request using a registered block leaves, the block is not freed
    but parked in an LRU of evictable blocks, so a later request with the same prefix (system prompt, few-shot header, chat history)
    re-pins it instead of recomputing the prefill.  Allocation takes free blocks first and evicts least-recently-used parked ones
    only when it must."""

    def __init__(self, num_blocks: int, enable_prefix_caching: bool = False):
        from collections import OrderedDict

        self.num_blocks = num_blocks
        self.enable_prefix_caching = enable_prefix_caching
        self.free: List[int] = list(range(num_blocks - 1, -1, -1))
        self.ref = [0] * num_blocks
        self.hash_of: Dict[int, int] = {}            # block -> prefix hash
        self.block_of: Dict[int, int] = {}           # prefix hash -> block
        self.parked = OrderedDict()                  # ref-zero registered blocks, oldest first
        self.hits = self.evictions = 0

    def allocate(self, n: int) -> Optional[List[int]]:
        if n > self.num_free:
            return None
        out = []
        for _ in range(n):
            if self.free:
                b = self.free.pop()
            else:
                b, _ = self.parked.popitem(last=False)
                self.block_of.pop(self.hash_of.pop(b), None)
                self.evictions += 1
            self.ref[b] = 1
            out.append(b)
        return out

    def release(self, blocks: List[int]):
        for b in reversed(blocks):
            self.ref[b] -= 1
            if self.ref[b] > 0:
                continue
            if b in self.hash_of:
                self.parked[b] = None
                self.parked.move_to_end(b)
            else:
                self.free.append(b)

    @property
    def num_free(self) -> int:
        return len(self.free) + len(self.parked)

    # ---- prefix cache ------------------------------------------------------------------------
    @staticmethod
    def chain_hashes(tokens: List[int], block_size: int) -> List[int]:
        """Hash of every FULL block, each folded over the previous one so equal hashes mean equal prefixes."""
        out, h = [], 0
        for i in range(0, len(tokens) - len(tokens) % block_size, block_size):
            h = hash((h, tuple(tokens[i : i + block_size])))
            out.append(h)
        return out

    def lookup_and_pin(self, hashes: List[int]) -> List[int]:
        """Longest run of leading hashes that are cached → their blocks, pinned for the caller."""
        got = []
        for h in hashes:
            b = self.block_of.get(h)
            if b is None:
                break
            if self.ref[b] == 0:
                self.parked.pop(b, None)
            self.ref[b] += 1
            got.append(b)
        self.hits += len(got)
        return got

    def register(self, block: int, h: int) -> None:
        if h in self.block_of or block in self.hash_of:
            return                                   # first writer wins; a concurrent duplicate stays private
        self.hash_of[block], self.block_of[h] = h, block


class PagedKVCache:
    """``k/v`` pools: [layers, num_blocks, block_size, kv_heads, head_dim]."""

    def __init__(self, num_layers: int, num_blocks: int, block_size: int, kv_heads: int, head_dim: int, dtype, device, enable_prefix_caching: bool = False):
        self.block_size = block_size
        self.enable_prefix_caching = enable_prefix_caching
        self.prefix_hit_tokens: Dict[int, int] = {}
        self.k = torch.zeros(num_layers, num_blocks, block_size, kv_heads, head_dim, dtype=dtype, device=device)
        self.v = torch.zeros_like(self.k)
        self.allocator = KVBlockAllocator(num_blocks, enable_prefix_caching)
        self.block_tables: Dict[int, List[int]] = {}
        self.lengths: Dict[int, int] = {}

    def scratch_block(self) -> int:
        """A block that belongs to no request (allocated on first use, never released): padding rows of a bucketed batch write here."""
        if getattr(self, "_scratch", None) is None:
            got = self.allocator.allocate(1)
            if got is None:
                raise RuntimeError("no free KV block left for the padding scratch block")
            self._scratch = got[0]
        return self._scratch

    def can_admit(self, num_tokens: int) -> bool:
        return self.allocator.num_free >= (num_tokens + self.block_size - 1) // self.block_size

    def add_request(self, rid: int, num_tokens: int, prompt_tokens: Optional[List[int]] = None) -> bool:
        """Reserve blocks for ``num_tokens``.  With prefix caching and the prompt given, leading blocks already in the cache are reused:
        ``prefix_hit_tokens[rid]`` says how many prompt tokens need no prefill (at least one token is always left to compute, because
        the logits of the last prompt position are needed)."""
        shared: List[int] = []
        if self.enable_prefix_caching and prompt_tokens is not None:
            hashes = self.allocator.chain_hashes(list(prompt_tokens), self.block_size)
            if len(hashes) * self.block_size == len(prompt_tokens):
                hashes = hashes[:-1]
            shared = self.allocator.lookup_and_pin(hashes)
        need = (num_tokens + self.block_size - 1) // self.block_size - len(shared)
        blocks = self.allocator.allocate(max(need, 0))
        if blocks is None:
            self.allocator.release(shared)
            return False
        self.block_tables[rid] = shared + blocks
        self.lengths[rid] = len(shared) * self.block_size
        self.prefix_hit_tokens[rid] = len(shared) * self.block_size
        return True

    def register_prefix(self, rid: int, prompt_tokens: List[int]) -> None:
        """After the prefill: publish this request's full prompt blocks for later requests."""
        if not self.enable_prefix_caching:
            return
        for b, h in zip(self.block_tables[rid], self.allocator.chain_hashes(list(prompt_tokens), self.block_size)):
            self.allocator.register(b, h)

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
        self.prefix_hit_tokens.pop(rid, None)

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

    def __init__(self, cache: PagedKVCache, rids: List[int], layer_numbers: List[int], pad_to: Optional[int] = None, max_len_multiple: int = 1):
        """``pad_to``: static batch size (a CUDA-graph bucket) — the extra rows are dummies at position 0 whose K/V land in the cache's scratch
        block and whose logits are dropped.  ``max_len_multiple`` rounds the attended length up so the gather shape changes rarely."""
        self.cache, self.rids = cache, rids
        self.block_table, self.lengths = cache.batch_tables(rids)
        self.num_real = len(rids)
        if pad_to is not None and pad_to > len(rids):
            extra = pad_to - len(rids)
            scratch = cache.scratch_block()
            self.block_table = torch.cat([self.block_table, torch.full((extra, self.block_table.shape[1]), scratch, dtype=torch.long, device=self.block_table.device)])
            self.lengths = torch.cat([self.lengths, self.lengths.new_zeros(extra)])
        self.max_len = int(max(cache.lengths[r] for r in rids)) + 1
        if max_len_multiple > 1:
            self.max_len = -(-self.max_len // max_len_multiple) * max_len_multiple
            need = -(-self.max_len // cache.block_size) - self.block_table.shape[1]
            if need > 0:                                             # columns past a request's blocks are masked out: any valid block id will do
                filler = torch.full((self.block_table.shape[0], need), cache.scratch_block(), dtype=torch.long, device=self.block_table.device)
                self.block_table = torch.cat([self.block_table, filler], dim=1)
        # int32 copies for the paged-attention kernels (made once per step, not once per layer)
        self.block_table_i32 = self.block_table.to(torch.int32).contiguous()
        self.positions_i32 = self.lengths.to(torch.int32).contiguous()
        self.lengths_incl_i32 = (self.lengths + 1).to(torch.int32).contiguous()
        self.max_sequence_length = self.max_len                      # rotary table length
        self.max_batch_size = len(rids)
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict: Dict = {}
        self.layer_index = {n: i for i, n in enumerate(layer_numbers)}


def _bdc_copy_from(self, other: "BatchedDecodeContext") -> None:
    """Refresh a (CUDA-graph static) context in place with another step's tables and positions; shapes must match (same bucket)."""
    assert self.block_table.shape == other.block_table.shape and self.lengths.shape == other.lengths.shape and self.max_len == other.max_len
    self.block_table.copy_(other.block_table)
    self.lengths.copy_(other.lengths)
    self.block_table_i32.copy_(other.block_table_i32)
    self.positions_i32.copy_(other.positions_i32)
    self.lengths_incl_i32.copy_(other.lengths_incl_i32)
    self.rids, self.num_real = other.rids, other.num_real


BatchedDecodeContext.copy_from = _bdc_copy_from


class PagedPrefillContext:
    """Inference context of a prefill WITHOUT cached prefix (understood by ``Attention.forward``): causal attention over the prompt itself, and each
    layer writes its rotated K / V for positions [0, n) straight into the request's pages."""

    is_paged_prefill = True

    def __init__(self, cache: PagedKVCache, rid: int, num_tokens: int, layer_numbers: List[int]):
        self.cache, self.rid, self.num_tokens = cache, rid, num_tokens
        self.layer_index = {n: i for i, n in enumerate(layer_numbers)}
        pos = torch.arange(num_tokens, device=cache.k.device)
        table = torch.tensor(cache.block_tables[rid], device=cache.k.device)
        self.blk, self.off = table[pos // cache.block_size], pos % cache.block_size
        self.max_sequence_length = num_tokens
        self.max_batch_size = 1
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict: Dict = {}

    def store(self, layer_number: int, key: torch.Tensor, value: torch.Tensor) -> None:
        """``key, value [n, 1, kv_heads, d]`` (already rotated)."""
        li = self.layer_index[layer_number]
        self.cache.k[li, self.blk, self.off] = key[:, 0]
        self.cache.v[li, self.blk, self.off] = value[:, 0]
