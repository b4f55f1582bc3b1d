"""Sharded parameter / gradient storage of Megatron-FSDP and the two communication pipelines on top of it
(reference ``distributed/fsdp/src/megatron_fsdp/param_and_grad_buffer.py``: ``DataParallelBuffer`` :1186, bucket allocators
:463-1163, ``ParamAndGradBuffer`` :1965, ``GradReducePipeline`` :3820, ``AllGatherPipeline`` :4278).

Layout decisions for B200:

* one **bucket per parameter group** (same FSDP unit, dtype, trainability, expert-ness); a Llama-70B layer at DP=8 is a ~0.8 GB
  bf16 bucket → ≈1 ms over NVLink 5, long enough to amortise launch latency and short enough to hide behind one layer of compute;
* bucket sizes are padded to ``world × 256 B`` so every rank's shard starts on a 256-byte boundary (TMA / 16-byte vector friendly,
  and legal for the NVLS ``multimem`` all-gather of ``parallel/nvlink.py`` when the pool lives in the symmetric heap);
* gathered weights live in a **fixed pool** of ``size`` (default 2) bucket groups — *double buffering*: unit i+1 is gathered into
  the other slot while unit i computes.  Fixed addresses are what CUDA graphs and registered (symmetric) communication buffers
  need; with 180 GB of HBM the pool is a rounding error, the point is the absence of allocator traffic and fragmentation;
* gradients are reduce-scattered per bucket as soon as the last gradient of the group has been produced, accumulated into an
  fp32 main-grad shard; at most ``suggested_queue_capacity`` bytes of reductions are in flight.
"""
from __future__ import annotations

import dataclasses
import math
from collections import defaultdict
from enum import Enum
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

ALIGN_BYTES = 256


def _pad(n: int, divisor: int) -> int:
    return (n + divisor - 1) // divisor * divisor


def _dtype_size(dtype: torch.dtype) -> int:
    return torch.empty((), dtype=dtype).element_size()


def _free_storage(t: torch.Tensor) -> None:
    if t.untyped_storage().size() > 0:
        t.untyped_storage().resize_(0)


def _alloc_storage(t: torch.Tensor, numel: int) -> None:
    need = numel * t.element_size()
    if t.untyped_storage().size() != need:
        t.untyped_storage().resize_(need)


# ====================================================================================================================
# policy / groups / index
# ====================================================================================================================
@dataclasses.dataclass
class BucketingPolicy:
    """``suggested_bucket_size`` (elements) chunks the parameters that do not belong to an FSDP unit; a unit is never split."""
    suggested_bucket_size: Optional[int] = 40_000_000
    fsdp_unit_modules: Sequence[type] = ()
    data_parallel_sharding_strategy: str = "no_shard"


@dataclasses.dataclass
class ParameterGroup:
    params: List[torch.nn.Parameter]
    dtype: Optional[torch.dtype] = None
    is_expert_param: bool = False
    requires_grad: Optional[bool] = None
    fsdp_unit_id: Optional[int] = None
    chunk_size_factor: int = 1
    model_weight_buffer: Optional["DataParallelBuffer"] = None
    main_weight_buffer: Optional["DataParallelBuffer"] = None
    main_grad_buffer: Optional["DataParallelBuffer"] = None


@dataclasses.dataclass
class BucketIndex:
    bucket_id: int
    global_data_index: int
    size: int
    items: List[int]


@dataclasses.dataclass
class ItemIndex:
    item_id: int
    global_data_index: int      # offset of the item in the (unsharded) bucket
    size: int
    shape: torch.Size


@dataclasses.dataclass
class ShardIndex:
    global_data_index: int      # offset of this rank's shard in the bucket
    local_data_index: int       # offset in the local buffer (0: one bucket per buffer)
    size: int


def build_data_parallel_buffer_index(elements: Sequence[torch.Size], data_parallel_rank: int, data_parallel_world_size: int, is_data_distributed: bool,
                                     ddp_config=None, bucket_id: int = 0, chunk_size_factor: int = 1, dtype: torch.dtype = torch.float32):
    """Item offsets (each item starts on a 16-byte boundary), the bucket padded to ``world × lcm(256 B, chunk)`` and this
    rank's shard of it.  Reference ``param_and_grad_buffer.py:258``."""
    esize = _dtype_size(dtype)
    item_align = max(1, 16 // esize)
    items, off = [], 0
    for i, shape in enumerate(elements):
        off = _pad(off, item_align)
        n = int(math.prod(shape)) if len(shape) else 1
        items.append(ItemIndex(i, off, n, torch.Size(shape)))
        off += n
    unit = math.lcm(max(1, ALIGN_BYTES // esize), max(1, chunk_size_factor))
    size = _pad(max(off, 1), data_parallel_world_size * unit)
    bucket = BucketIndex(bucket_id, 0, size, [it.item_id for it in items])
    shard_size = size // data_parallel_world_size
    if is_data_distributed:
        shard = ShardIndex(data_parallel_rank * shard_size, 0, shard_size)
    else:
        shard = ShardIndex(0, 0, size)
    return items, bucket, shard


@dataclasses.dataclass
class Bucket:
    data: torch.Tensor
    bucket_id: int = -1


# ====================================================================================================================
# temporary-bucket allocators
# ====================================================================================================================
class TemporaryBucketAllocator:
    """Fresh tensor per request; the caching allocator does the recycling.  Simple, but addresses change every step."""

    def __init__(self):
        self.buckets: Dict[int, Bucket] = {}

    def allocate(self, bucket_id: int, size: int, dtype: torch.dtype, device, mem_alloc_context: Optional[Callable] = None) -> Bucket:
        b = self.buckets.get(bucket_id)
        if b is None:
            b = Bucket(torch.empty(size, dtype=dtype, device=device), bucket_id)
            self.buckets[bucket_id] = b
        return b

    def free(self, bucket_id: int):
        self.buckets.pop(bucket_id, None)

    def can_allocate(self, bucket_id: int) -> bool:
        """Would ``allocate`` succeed WITHOUT leaving the allocator's preferred (pooled) path?  Prefetchers ask first."""
        return True


class StorageResizeBasedBucketAllocator(TemporaryBucketAllocator):
    """One persistent tensor object per bucket whose STORAGE comes and goes: views handed out earlier stay valid objects."""

    def allocate(self, bucket_id, size, dtype, device, mem_alloc_context=None) -> Bucket:
        b = self.buckets.get(bucket_id)
        if b is None:
            b = Bucket(torch.empty(size, dtype=dtype, device=device), bucket_id)
            self.buckets[bucket_id] = b
        _alloc_storage(b.data, size)
        return b

    def free(self, bucket_id: int):
        b = self.buckets.get(bucket_id)
        if b is not None:
            _free_storage(b.data)


class RotaryBucketAllocator(TemporaryBucketAllocator):
    """A ring of named buffers: a request takes the lowest idle slot (growing it if the bucket is larger than anything the
    slot has held); ``free`` returns the slot.  Memory high-water mark = (#buckets simultaneously alive) × (largest bucket)."""

    def __init__(self, name: str = "rotary"):
        super().__init__()
        self.name = name
        self.slots: List[Optional[torch.Tensor]] = []
        self.idle: List[int] = []
        self.using: Dict[int, int] = {}

    def allocate(self, bucket_id, size, dtype, device, mem_alloc_context=None) -> Bucket:
        if bucket_id in self.using:
            return self.buckets[bucket_id]
        if self.idle:
            slot = min(self.idle)
            self.idle.remove(slot)
        else:
            slot = len(self.slots)
            self.slots.append(None)
        nbytes = size * _dtype_size(dtype)
        raw = self.slots[slot]
        if raw is None or raw.numel() < nbytes or raw.device != torch.device(device):
            raw = torch.empty(_pad(nbytes, ALIGN_BYTES), dtype=torch.uint8, device=device)
            self.slots[slot] = raw
        self.using[bucket_id] = slot
        b = Bucket(raw[:nbytes].view(dtype), bucket_id)
        self.buckets[bucket_id] = b
        return b

    def _get_gbuf_name(self, buffer_id: int) -> str:
        return f"{self.name}_{buffer_id}"

    def free(self, bucket_id: int):
        slot = self.using.pop(bucket_id, None)
        if slot is not None:
            self.idle.append(slot)
            self.buckets.pop(bucket_id, None)


class FixedPoolAllocator(TemporaryBucketAllocator):
    """``size`` buffer groups shaped like the most common FSDP unit (the transformer layer): the k-th such unit ALWAYS uses group
    ``k mod size``.  The fixed assignment is a correctness requirement, not a nicety: autograd saves VIEWS of the gathered weights
    (``weight.t()`` in ``F.linear``) that alias the bucket memory of the forward pass, so the backward re-gather must land in the
    same memory.  Buckets of units with another shape (embedding, output layer) use per-bucket storage-resize allocation, or stay
    resident when ``fallback_to_persistent_buffer``.  ``strict`` (weights): a busy group at demand time raises instead of silently
    gathering elsewhere; gradient buckets are never aliased by saved tensors and simply take the slow path."""

    def __init__(self, name: str, fsdp_param_groups: List[ParameterGroup], size: int = 2, dtype_fn: Callable[[ParameterGroup], torch.dtype] = lambda g: g.dtype,
                 fallback_to_persistent_buffer: bool = False, strict: bool = False):
        super().__init__()
        self.name, self.groups, self.size, self.dtype_fn, self.strict = name, fsdp_param_groups, size, dtype_fn, strict
        unit_buckets: Dict[int, List[int]] = defaultdict(list)
        for bid, g in enumerate(fsdp_param_groups):
            if g.fsdp_unit_id is not None:
                unit_buckets[g.fsdp_unit_id].append(bid)
        self.fsdp_unit_buckets = unit_buckets
        # the largest family of units with identical bucket signatures shares the pool
        best: List[int] = []
        for u, bids in unit_buckets.items():
            fam = [v for v, other in unit_buckets.items() if self._is_two_bucket_group_equal(other, bids)]
            if len(fam) > len(best):
                best = fam
        assert best, "FixedPoolAllocator: no FSDP units to double-buffer"
        self.fsdp_double_buffer_units = best
        self.bucket_offset = {bid: k for u in best for k, bid in enumerate(unit_buckets[u])}
        self.bucket_group = {bid: order % size for order, u in enumerate(sorted(best)) for bid in unit_buckets[u]}
        n_per_group = len(unit_buckets[best[0]])
        self.idle_buffer: List[Tuple[int, int]] = [(g, k) for g in range(size) for k in range(n_per_group)]
        self.using_buffer: Dict[int, Tuple[int, int]] = {}
        self.pool: Dict[Tuple[int, int], torch.Tensor] = {}
        self.fallback_to_persistent_buffer = fallback_to_persistent_buffer
        self.backup_allocator = StorageResizeBasedBucketAllocator()
        self.pool_misses = 0

    def _sig(self, bid: int):
        g = self.groups[bid]
        return (self.dtype_fn(g), sum(p.numel() for p in g.params))

    def _is_two_bucket_group_equal(self, group_a: Iterable[int], group_b: Iterable[int]) -> bool:
        a, b = list(group_a), list(group_b)
        return len(a) == len(b) and all(self._sig(x) == self._sig(y) for x, y in zip(a, b))

    def _get_gbuf_name(self, buf_group_id: int, bucket_index: int) -> str:
        return f"{self.name}_{buf_group_id}_{bucket_index}"

    def allocate(self, bucket_id, size, dtype, device, mem_alloc_context=None) -> Bucket:
        if bucket_id in self.buckets:                        # live allocation (pooled OR fallback): the same memory until free()
            return self.buckets[bucket_id]
        k = self.bucket_offset.get(bucket_id)
        if k is not None:
            for slot in self.idle_buffer:
                if slot == (self.bucket_group[bucket_id], k):
                    self.idle_buffer.remove(slot)
                    self.using_buffer[bucket_id] = slot
                    raw = self.pool.get(slot)
                    nbytes = size * _dtype_size(dtype)
                    if raw is None or raw.numel() < nbytes:
                        raw = torch.empty(_pad(nbytes, ALIGN_BYTES), dtype=torch.uint8, device=device)
                        self.pool[slot] = raw
                    b = Bucket(raw[:nbytes].view(dtype), bucket_id)
                    self.buckets[bucket_id] = b
                    return b
            if self.strict:
                holder = [b for b, sl in self.using_buffer.items() if sl == (self.bucket_group[bucket_id], k)]
                raise RuntimeError(f"{self.name}: bucket {bucket_id} needs buffer group {self.bucket_group[bucket_id]}, still held by bucket(s) {holder}: "
                                   f"raise the double-buffer depth (size={self.size}) or release the holder first")
            self.pool_misses += 1                            # (gradient buckets) take the slow path
        b = self.backup_allocator.allocate(bucket_id, size, dtype, device)
        self.buckets[bucket_id] = b
        return b

    def can_allocate(self, bucket_id: int) -> bool:
        k = self.bucket_offset.get(bucket_id)
        return k is None or bucket_id in self.buckets or (self.bucket_group[bucket_id], k) in self.idle_buffer

    def free(self, bucket_id: int):
        slot = self.using_buffer.pop(bucket_id, None)
        if slot is not None:
            self.idle_buffer.append(slot)
            self.buckets.pop(bucket_id, None)
        else:
            self.buckets.pop(bucket_id, None)
            if not self.fallback_to_persistent_buffer:
                self.backup_allocator.free(bucket_id)


class MaxPoolAllocator(TemporaryBucketAllocator):
    """Double buffering for HETEROGENEOUS units (hybrid Mamba / attention / MoE stacks): per dtype the pool holds, for the j-th
    largest bucket of any unit, a buffer of the maximum size over units; a unit's buckets are matched to pool entries by size
    rank.  Every unit fits, the pool is ``size`` × (per-rank maxima)."""

    def __init__(self, name: str, fsdp_param_groups: List[ParameterGroup], size: int = 2, dtype_fn: Callable[[ParameterGroup], torch.dtype] = lambda g: g.dtype,
                 fallback_to_persistent_buffer: bool = False, strict: bool = False):
        super().__init__()
        self.name, self.groups, self.size, self.dtype_fn, self.strict = name, fsdp_param_groups, size, dtype_fn, strict
        self.fsdp_unit_buckets: Dict[int, List[int]] = defaultdict(list)
        for bid, g in enumerate(fsdp_param_groups):
            if g.fsdp_unit_id is not None:
                self.fsdp_unit_buckets[g.fsdp_unit_id].append(bid)
        assert self.fsdp_unit_buckets, "MaxPoolAllocator: no FSDP units to double-buffer"
        self.fsdp_double_buffer_units = list(self.fsdp_unit_buckets)
        self.bucket_group = {bid: order % size for order, u in enumerate(sorted(self.fsdp_unit_buckets)) for bid in self.fsdp_unit_buckets[u]}
        self.max_dtype_bucket_sizes: Dict[torch.dtype, List[int]] = {}
        self.bucket_alloc_index: Dict[int, Tuple[torch.dtype, int]] = {}
        self._build_fixed_max_pool()
        self.idle_buffer = [(g, dt, k) for g in range(size) for dt, sizes in self.max_dtype_bucket_sizes.items() for k in range(len(sizes))]
        self.using_buffer: Dict[int, Tuple[int, torch.dtype, int]] = {}
        self.pool: Dict[Tuple[int, torch.dtype, int], torch.Tensor] = {}
        self.fallback_to_persistent_buffer = fallback_to_persistent_buffer
        self.backup_allocator = StorageResizeBasedBucketAllocator()
        self.pool_misses = 0

    def _numel(self, bid: int) -> int:
        return sum(p.numel() for p in self.groups[bid].params)

    def _build_fixed_max_pool(self):
        for u, bids in self.fsdp_unit_buckets.items():
            per_dtype: Dict[torch.dtype, List[int]] = defaultdict(list)
            for bid in bids:
                per_dtype[self.dtype_fn(self.groups[bid])].append(bid)
            for dt, lst in per_dtype.items():
                lst.sort(key=self._numel)                       # smallest bucket <-> smallest pool entry
                sizes = self.max_dtype_bucket_sizes.setdefault(dt, [])
                for k, bid in enumerate(lst):
                    if k == len(sizes):
                        sizes.append(0)
                    self.bucket_alloc_index[bid] = (dt, k)
                # ranks are aligned from the top so that the largest buckets share the largest entry
            # second pass below fixes the maxima once all units are known
        for dt, sizes in self.max_dtype_bucket_sizes.items():
            for k in range(len(sizes)):
                sizes[k] = max((self._numel(b) for b, (d, kk) in self.bucket_alloc_index.items() if d == dt and kk == k), default=0)

    def _get_gbuf_name(self, buf_group_id: int, dtype: torch.dtype, bucket_index: int) -> str:
        return f"{self.name}_{buf_group_id}_{str(dtype).split('.')[-1]}_{bucket_index}"

    def allocate(self, bucket_id, size, dtype, device, mem_alloc_context=None) -> Bucket:
        if bucket_id in self.buckets:                        # live allocation (pooled OR fallback): the same memory until free()
            return self.buckets[bucket_id]
        idx = self.bucket_alloc_index.get(bucket_id)
        if idx is not None:
            for slot in self.idle_buffer:
                if slot == (self.bucket_group[bucket_id],) + idx:
                    self.idle_buffer.remove(slot)
                    self.using_buffer[bucket_id] = slot
                    raw = self.pool.get(slot)
                    # the padded bucket may exceed the raw parameter count: size the entry for the largest PADDED request seen
                    if raw is None or raw.numel() < size:
                        raw = torch.empty(max(size, self.max_dtype_bucket_sizes[idx[0]][idx[1]]), dtype=dtype, device=device)
                        self.pool[slot] = raw
                    b = Bucket(raw[:size], bucket_id)
                    self.buckets[bucket_id] = b
                    return b
            if self.strict:
                raise RuntimeError(f"{self.name}: bucket {bucket_id} needs buffer group {self.bucket_group[bucket_id]} which is still in use "
                                   f"(in use: {sorted(self.using_buffer)}); raise the double-buffer depth (size={self.size})")
            self.pool_misses += 1
        b = self.backup_allocator.allocate(bucket_id, size, dtype, device)
        self.buckets[bucket_id] = b
        return b

    def can_allocate(self, bucket_id: int) -> bool:
        idx = self.bucket_alloc_index.get(bucket_id)
        return idx is None or bucket_id in self.buckets or ((self.bucket_group[bucket_id],) + idx) in self.idle_buffer

    def free(self, bucket_id: int):
        slot = self.using_buffer.pop(bucket_id, None)
        if slot is not None:
            self.idle_buffer.append(slot)
            self.buckets.pop(bucket_id, None)
        else:
            self.buckets.pop(bucket_id, None)
            if not self.fallback_to_persistent_buffer:
                self.backup_allocator.free(bucket_id)


# ====================================================================================================================
# one flat (possibly sharded) buffer of one parameter group
# ====================================================================================================================
class DataParallelBuffer:
    """Flat storage of ONE parameter group in ONE role (model weights / fp32 main weights / main grads).

    ``is_data_distributed``: only this rank's 1/world slice of the bucket is resident (``self.data``); the full bucket exists
    temporarily (``fetch_bucket`` / ``allocate_bucket_storage``) in memory owned by ``temporary_bucket_allocator``."""

    def __init__(self, ddp_config, params: List[torch.nn.Parameter], is_data_distributed: bool, bucket_id: int, dtype: Optional[torch.dtype] = None,
                 device=None, data_parallel_group=None, temporary_bucket_allocator: Optional[TemporaryBucketAllocator] = None, is_dtype_float8: bool = False,
                 gradient_scaling_factor: Optional[float] = None, chunk_size_factor: int = 1, mem_alloc_context: Optional[Callable] = None,
                 index_dtype: Optional[torch.dtype] = None):
        """``index_dtype``: the dtype whose alignment rules lay the bucket out — the three buffers of a group (bf16 weights, fp32
        main weights, fp32 main grads) must agree element for element on item offsets and shard boundaries."""
        self.ddp_config, self.params = ddp_config, list(params)
        self.is_data_distributed, self.bucket_id = is_data_distributed, bucket_id
        self.dtype = dtype if dtype is not None else params[0].dtype
        self.device = device if device is not None else params[0].device
        self.data_parallel_group = data_parallel_group
        self.dp_rank = dist.get_rank(data_parallel_group) if dist.is_initialized() else 0
        self.dp_world_size = dist.get_world_size(data_parallel_group) if dist.is_initialized() else 1
        self.temporary_bucket_allocator = temporary_bucket_allocator if temporary_bucket_allocator is not None else TemporaryBucketAllocator()
        self.gradient_scaling_factor = gradient_scaling_factor
        self.param_idx = {id(p): i for i, p in enumerate(self.params)}
        self.item_index_map, self.bucket_index, self.shard_bucket_index = build_data_parallel_buffer_index(
            [p.shape for p in self.params], self.dp_rank, self.dp_world_size, is_data_distributed, ddp_config, bucket_id, chunk_size_factor,
            index_dtype if index_dtype is not None else self.dtype)
        self.data_size = self.shard_bucket_index.size
        self.data: Optional[torch.Tensor] = None

    # ---- storage ---------------------------------------------------------------------------------------------------
    def init_data(self, data: torch.Tensor):
        assert data.numel() == self.data_size and data.dtype == self.dtype, (data.numel(), self.data_size, data.dtype, self.dtype)
        self.data = data

    def fetch_bucket(self, dtype: Optional[torch.dtype] = None, set_param_data: bool = False) -> Bucket:
        """The full bucket: the resident buffer itself if not distributed, else a temporary one (contents undefined until an
        all-gather fills it).  ``set_param_data`` re-points the parameters at the bucket."""
        if not self.is_data_distributed:
            b = Bucket(self.data, self.bucket_id)
        else:
            b = self.temporary_bucket_allocator.allocate(self.bucket_id, self.bucket_index.size, dtype or self.dtype, self.device)
        if set_param_data:
            for p in self.params:
                p.data = self.get_item_from_bucket(b, self.param_idx[id(p)]).view(p.shape if p.numel() else self.item_index_map[self.param_idx[id(p)]].shape)
        return b

    def allocate_bucket_storage(self, dtype: Optional[torch.dtype] = None) -> Bucket:
        return self.fetch_bucket(dtype)

    def free_bucket_storage(self):
        if self.is_data_distributed:
            self.temporary_bucket_allocator.free(self.bucket_id)

    # ---- index math ----------------------------------------------------------------------------------------------------
    def _get_item_slice_in_shard(self, item_id: int) -> Tuple[int, int]:
        """[start, end) of the part of item ``item_id`` that falls into this rank's shard, relative to the ITEM."""
        it, sh = self.item_index_map[item_id], self.shard_bucket_index
        lo = max(it.global_data_index, sh.global_data_index)
        hi = min(it.global_data_index + it.size, sh.global_data_index + sh.size)
        if lo >= hi:
            return 0, 0
        return lo - it.global_data_index, hi - it.global_data_index

    def locate_item_in_global_item(self, item_id: int) -> Tuple[int, int]:
        return self._get_item_slice_in_shard(item_id)

    def _get_item_local_shard_index(self, item_id: int) -> Tuple[int, int]:
        """[start, end) of that part inside the local shard buffer."""
        s, e = self._get_item_slice_in_shard(item_id)
        if s == e:
            return 0, 0
        base = self.item_index_map[item_id].global_data_index - self.shard_bucket_index.global_data_index + self.shard_bucket_index.local_data_index
        return base + s, base + e

    def _get_item_local_index(self, item_id: int) -> Tuple[int, int]:
        if self.is_data_distributed:
            return self._get_item_local_shard_index(item_id)
        it = self.item_index_map[item_id]
        return it.global_data_index, it.global_data_index + it.size

    # ---- items -----------------------------------------------------------------------------------------------------------
    def set_item(self, item_id: int, item_data: torch.Tensor) -> None:
        """Copy a FULL item in; a distributed buffer keeps only its slice."""
        if self.is_data_distributed:
            s, e = self._get_item_slice_in_shard(item_id)
            if s == e:
                return
            ls, le = self._get_item_local_shard_index(item_id)
            self.data[ls:le].copy_(item_data.detach().reshape(-1)[s:e])
        else:
            ls, le = self._get_item_local_index(item_id)
            self.data[ls:le].copy_(item_data.detach().reshape(-1))

    def get_item(self, item_id: int, only_shard: bool = False) -> torch.Tensor:
        if only_shard or self.is_data_distributed:
            ls, le = self._get_item_local_shard_index(item_id) if self.is_data_distributed else self._shard_of_full(item_id)
            return self.data[ls:le]
        ls, le = self._get_item_local_index(item_id)
        return self.data[ls:le].view(self.item_index_map[item_id].shape)

    def _shard_of_full(self, item_id: int) -> Tuple[int, int]:
        """Slice of a NON-distributed buffer that the same rank would own if it were distributed (ZeRO-1 optimizer shards)."""
        it = self.item_index_map[item_id]
        ssz = self.bucket_index.size // self.dp_world_size
        lo = max(it.global_data_index, self.dp_rank * ssz)
        hi = min(it.global_data_index + it.size, (self.dp_rank + 1) * ssz)
        return (lo, hi) if lo < hi else (0, 0)

    def get_item_from_bucket(self, bucket: Bucket, item_id: int) -> torch.Tensor:
        it = self.item_index_map[item_id]
        return bucket.data[it.global_data_index: it.global_data_index + it.size]

    def get_shard_from_bucket(self, bucket: Bucket) -> torch.Tensor:
        ssz = self.bucket_index.size // self.dp_world_size
        return bucket.data[self.dp_rank * ssz: (self.dp_rank + 1) * ssz]

    def get_shard_from_local_buffer(self) -> torch.Tensor:
        if self.is_data_distributed:
            return self.data[self.shard_bucket_index.local_data_index: self.shard_bucket_index.local_data_index + self.shard_bucket_index.size]
        return self.get_shard_from_bucket(Bucket(self.data, self.bucket_id))


# ====================================================================================================================
# grouping
# ====================================================================================================================
def _is_expert_param(p) -> bool:
    return not getattr(p, "allreduce", True)


def _get_parameter_groups(module: torch.nn.Module, policy: BucketingPolicy, meta_device_init_fp8_params: Optional[dict] = None,
                          bucket_group_by_fsdp_unit: bool = True) -> Tuple[List[ParameterGroup], Dict[int, int], Dict[int, List[int]]]:
    """Parameters -> groups (= buckets).  Returns (groups, id(param) -> group index, fsdp unit id -> bucket ids in gather order).
    Reference ``param_and_grad_buffer.py:1731``."""
    unit_of: Dict[int, int] = {}
    n_units = 0
    if policy.fsdp_unit_modules:
        for m in module.modules():
            if isinstance(m, tuple(policy.fsdp_unit_modules)):
                fresh = [p for p in m.parameters() if id(p) not in unit_of]
                if fresh:
                    for p in fresh:
                        unit_of[id(p)] = n_units
                    n_units += 1
    keyed: Dict[Tuple, List[torch.nn.Parameter]] = {}
    seen = set()
    for p in module.parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        key = (unit_of.get(id(p), -1), p.dtype, p.requires_grad, _is_expert_param(p))
        keyed.setdefault(key, []).append(p)
    groups: List[ParameterGroup] = []
    for (unit, dtype, rg, expert), params in keyed.items():
        if unit >= 0 or not policy.suggested_bucket_size:
            groups.append(ParameterGroup(params, dtype, expert, rg, unit if unit >= 0 else None))
            continue
        cur, n = [], 0
        for p in params:
            cur.append(p)
            n += p.numel()
            if n >= policy.suggested_bucket_size:
                groups.append(ParameterGroup(cur, dtype, expert, rg, None))
                cur, n = [], 0
        if cur:
            groups.append(ParameterGroup(cur, dtype, expert, rg, None))
    param_to_group = {id(p): gi for gi, g in enumerate(groups) for p in g.params}
    unit_buckets: Dict[int, List[int]] = defaultdict(list)
    for gi, g in enumerate(groups):
        if g.fsdp_unit_id is not None:
            unit_buckets[g.fsdp_unit_id].append(gi)
    return groups, param_to_group, dict(unit_buckets)


def gradient_reduce_preprocessing(grad_data: torch.Tensor, scaling_factor: Optional[float], ddp_config) -> dist.ReduceOp:
    """Pre-scale (or pick AVG) so that the collective yields the data-parallel mean (reference :4740)."""
    if scaling_factor is None or scaling_factor == 1.0:
        return dist.ReduceOp.SUM
    if getattr(ddp_config, "average_in_collective", False) and grad_data.is_cuda:
        return dist.ReduceOp.AVG
    grad_data.mul_(scaling_factor)
    return dist.ReduceOp.SUM


def _check_nan_in_grad(grad: torch.Tensor, name: str = ""):
    if not torch.isfinite(grad).all():
        raise RuntimeError(f"non-finite gradient in bucket {name} before the data-parallel reduction")


# ====================================================================================================================
# the container
# ====================================================================================================================
_SHARD_WEIGHTS = ("optim_grads_params",)
_SHARD_GRADS = ("optim_grads", "optim_grads_params")
_SHARD_OPTIM = ("optim", "optim_grads", "optim_grads_params")


class ParamAndGradBuffer:
    """All parameter groups of one module with their three buffers, plus the optimizer's view of them.

    ``optimizer_named_parameters``: fp32 (``preserve_fp32_weights``) parameters over this rank's SHARD of every group — the only
    tensors the optimizer ever sees; their ``.grad`` are views of the main-grad shards (``update_main_grads``)."""

    def __init__(self, ddp_config, module: torch.nn.Module, bucketing_policy: BucketingPolicy, data_parallel_group=None, expert_data_parallel_group=None,
                 preserve_fp32_weights: bool = True, grad_reduce_in_fp32: bool = True, gradient_scaling_factor: Optional[float] = None,
                 expert_gradient_scaling_factor: Optional[float] = None, device=None, reset_parameters_for_meta_device_init_module: bool = False,
                 allocator: str = "auto", double_buffer_size: int = 2):
        self.ddp_config, self.module, self.bucketing_policy = ddp_config, module, bucketing_policy
        self.strategy = bucketing_policy.data_parallel_sharding_strategy
        assert self.strategy in ("no_shard", "optim", "optim_grads", "optim_grads_params"), self.strategy
        self.dp_group, self.expert_dp_group = data_parallel_group, expert_data_parallel_group or data_parallel_group
        self.preserve_fp32_weights, self.grad_reduce_in_fp32 = preserve_fp32_weights, grad_reduce_in_fp32
        self.gradient_scaling_factor, self.expert_gradient_scaling_factor = gradient_scaling_factor, expert_gradient_scaling_factor
        self.parameter_groups, self.param_to_param_group, self.fsdp_unit_buckets = _get_parameter_groups(module, bucketing_policy)
        self.param_to_name = {id(p): n for n, p in module.named_parameters()}
        self.device = device
        shard_w, shard_g = self.strategy in _SHARD_WEIGHTS, self.strategy in _SHARD_GRADS
        has_units = any(g.fsdp_unit_id is not None for g in self.parameter_groups)
        if allocator == "auto":
            allocator = "fixed" if (getattr(ddp_config, "fsdp_double_buffer", False) and has_units) else "resize"
        mk = {
            "temporary": lambda n, dt, st: TemporaryBucketAllocator(), "resize": lambda n, dt, st: StorageResizeBasedBucketAllocator(),
            "rotary": lambda n, dt, st: RotaryBucketAllocator(n),
            "fixed": lambda n, dt, st: FixedPoolAllocator(n, self.parameter_groups, double_buffer_size, dt, strict=st),
            "max": lambda n, dt, st: MaxPoolAllocator(n, self.parameter_groups, double_buffer_size, dt, strict=st),
        }[allocator]
        self.weight_alloc = mk("fsdp_params", lambda g: g.dtype, True)
        self.grad_alloc = mk("fsdp_grads", lambda g: torch.float32 if grad_reduce_in_fp32 else g.dtype, False)
        for gi, g in enumerate(self.parameter_groups):
            grp = self.expert_dp_group if g.is_expert_param else self.dp_group
            scale = expert_gradient_scaling_factor if g.is_expert_param else gradient_scaling_factor
            dev = device if device is not None else g.params[0].device
            g.model_weight_buffer = DataParallelBuffer(ddp_config, g.params, shard_w, gi, g.dtype, dev, grp, self.weight_alloc)
            full = torch.zeros(g.model_weight_buffer.bucket_index.size, dtype=g.dtype, device=dev)
            for i, p in enumerate(g.params):
                it = g.model_weight_buffer.item_index_map[i]
                full[it.global_data_index: it.global_data_index + it.size].copy_(p.detach().reshape(-1))
            if shard_w:
                g.model_weight_buffer.init_data(g.model_weight_buffer.get_shard_from_bucket(Bucket(full)).clone())
            else:
                g.model_weight_buffer.init_data(full)
            if g.requires_grad:
                need_main = preserve_fp32_weights and g.dtype != torch.float32
                if need_main or self.strategy in _SHARD_OPTIM:
                    # main weights always live on the optimizer shard only
                    g.main_weight_buffer = DataParallelBuffer(ddp_config, g.params, True, gi, torch.float32 if preserve_fp32_weights else g.dtype, dev, grp,
                                                              index_dtype=g.dtype)
                    mw = g.main_weight_buffer
                    mw.init_data(mw.get_shard_from_bucket(Bucket(full)).to(mw.dtype).clone())
                gdt = torch.float32 if grad_reduce_in_fp32 else g.dtype
                g.main_grad_buffer = DataParallelBuffer(ddp_config, g.params, shard_g, gi, gdt, dev, grp, self.grad_alloc, gradient_scaling_factor=scale,
                                                        index_dtype=g.dtype)
                g.main_grad_buffer.init_data(torch.zeros(g.main_grad_buffer.data_size, dtype=gdt, device=dev))
            # parameters become views of the weight storage (resident strategies) or are released (ZeRO-3)
            if shard_w:
                del full
                for p in g.params:
                    p._fsdp_shape = p.shape
                    p.data = torch.empty(0, dtype=p.dtype, device=dev)
            else:
                g.model_weight_buffer.fetch_bucket(set_param_data=True)
        self._init_optimizer_named_parameters()

    # ---- optimizer view -------------------------------------------------------------------------------------------------
    def _init_optimizer_named_parameters(self):
        self.optimizer_named_parameters: List[Tuple[str, torch.nn.Parameter]] = []
        self._opt_params: List[Tuple[ParameterGroup, torch.nn.Parameter]] = []
        for gi, g in enumerate(self.parameter_groups):
            if not g.requires_grad:
                continue
            if g.main_weight_buffer is not None:
                shard = g.main_weight_buffer.data
            else:
                shard = g.model_weight_buffer.data             # no_shard + fp32 params: the optimizer updates the weights in place
            q = torch.nn.Parameter(shard, requires_grad=True)
            q.fsdp_group_index = gi
            q.is_expert_param = g.is_expert_param
            self.optimizer_named_parameters.append((f"fsdp_group_{gi}", q))
            self._opt_params.append((g, q))

    def optimizer_parameters(self) -> List[torch.nn.Parameter]:
        return [q for _, q in self.optimizer_named_parameters]

    @property
    def num_buckets(self) -> int:
        return len(self.parameter_groups)

    def update_main_grads(self):
        """Point the optimizer parameters' ``.grad`` at the reduced gradient shards (reference :3351)."""
        for g, q in self._opt_params:
            gb = g.main_grad_buffer
            grad = gb.get_shard_from_local_buffer() if (gb.is_data_distributed or g.main_weight_buffer is not None) else gb.data
            q.grad = grad if grad.dtype == q.dtype else grad.to(q.dtype)

    def zero_grad(self):
        for g in self.parameter_groups:
            if g.main_grad_buffer is not None:
                g.main_grad_buffer.data.zero_()
            for p in g.params:
                p.grad = None
        for _, q in self._opt_params:
            q.grad = None

    def scale_gradients(self, scaling_factor: float) -> None:
        for g in self.parameter_groups:
            if g.main_grad_buffer is not None:
                g.main_grad_buffer.data.mul_(scaling_factor)

    @torch.no_grad()
    def copy_main_weights_to_model_weights(self):
        """After ``optimizer.step()``: cast the updated fp32 shards into the model-weight storage (reference :3418).  Resident
        strategies then need ``all_gather_parameters`` to refresh the other ranks' slices."""
        for g in self.parameter_groups:
            mw, w = g.main_weight_buffer, g.model_weight_buffer
            if mw is None:
                continue
            dst = w.data if w.is_data_distributed else w.get_shard_from_bucket(Bucket(w.data))
            dst.copy_(mw.data)

    def all_gather_parameters(self, async_op: bool = False):
        """Resident strategies: every rank's updated slice -> everybody (in place, the bucket IS the resident buffer)."""
        handles = []
        for g in self.parameter_groups:
            w = g.model_weight_buffer
            if w.is_data_distributed or g.main_weight_buffer is None or w.dp_world_size == 1:
                continue
            shard = w.get_shard_from_bucket(Bucket(w.data))
            handles.append(dist.all_gather_into_tensor(w.data, shard.clone() if not shard.is_cuda else shard, group=w.data_parallel_group, async_op=async_op))
        return [h for h in handles if h is not None]

    def _reduce_group(self, g: ParameterGroup, scatter: bool, async_op: bool = False):
        gb = g.main_grad_buffer
        if gb is None or gb.dp_world_size == 1:
            return None
        op = gradient_reduce_preprocessing(gb.data, gb.gradient_scaling_factor, self.ddp_config)
        if scatter:
            shard = gb.get_shard_from_bucket(Bucket(gb.data))
            out = torch.empty_like(shard)
            h = dist.reduce_scatter_tensor(out, gb.data, op=op, group=gb.data_parallel_group, async_op=async_op)
            if h is not None:
                h.wait()
            shard.copy_(out)
            return None
        return dist.all_reduce(gb.data, op=op, group=gb.data_parallel_group, async_op=async_op)

    def reduce_scatter_gradients(self, async_op: bool = False):
        """ZeRO-1: the resident full gradient buffers are reduce-scattered (every rank keeps the mean of ITS slice)."""
        for g in self.parameter_groups:
            if g.main_grad_buffer is not None and not g.main_grad_buffer.is_data_distributed:
                self._reduce_group(g, scatter=True, async_op=False)

    def all_reduce_gradients(self, async_op: bool = False):
        hs = [self._reduce_group(g, scatter=False, async_op=async_op) for g in self.parameter_groups
              if g.main_grad_buffer is not None and not g.main_grad_buffer.is_data_distributed]
        return [h for h in hs if h is not None]


# ====================================================================================================================
# pipelines
# ====================================================================================================================
class BucketStatus(Enum):
    EMPTY = 1
    COMMUNICATING = 2
    READY_TO_USE = 3


class PrefetchOrder(Enum):
    FORWARD_PASS_ORDER = 0
    BACKWARD_PASS_ORDER = 1


class AllGatherPipeline:
    """Asynchronous parameter un-sharding with look-ahead (reference :4278).

    ``all_gather_params(params, prefetch=True, prefetch_order=…, suggested_AG_prefetch_size=N)`` launches the gathers of the
    buckets holding ``params`` and keeps launching the following buckets (in forward or backward order) until ``N`` elements
    are in flight.  ``wait_bucket_ready`` blocks on the handle and re-points the parameters; ``release_bucket`` returns the
    bucket to the allocator."""

    def __init__(self, param_and_grad_buffer: ParamAndGradBuffer, async_op: bool = True):
        self.buffer, self.async_op = param_and_grad_buffer, async_op
        self.status = {i: BucketStatus.EMPTY for i in range(self.num_buckets)}
        self.handles: Dict[int, Optional[object]] = {}
        self.launch_log: List[int] = []                     # order of launches (tests / debugging)

    @property
    def num_buckets(self) -> int:
        return self.buffer.num_buckets

    def get_fsdp_buffer(self, bucket_id: int) -> DataParallelBuffer:
        return self.buffer.parameter_groups[bucket_id].model_weight_buffer

    def reset(self):
        for b in list(self.status):
            if self.status[b] != BucketStatus.EMPTY:
                self.release_bucket(b)

    def async_bucket_gather(self, bucket_id: int) -> None:
        if self.status[bucket_id] != BucketStatus.EMPTY:
            return
        w = self.get_fsdp_buffer(bucket_id)
        if not w.is_data_distributed:
            self.status[bucket_id] = BucketStatus.READY_TO_USE
            return
        bucket = w.fetch_bucket()
        if w.dp_world_size > 1:
            self.handles[bucket_id] = dist.all_gather_into_tensor(bucket.data, w.data, group=w.data_parallel_group, async_op=self.async_op)
        else:
            bucket.data.copy_(w.data)
            self.handles[bucket_id] = None
        self.status[bucket_id] = BucketStatus.COMMUNICATING
        self.launch_log.append(bucket_id)

    def all_gather_params(self, params: Sequence[torch.nn.Parameter], prefetch: bool = False, prefetch_order: PrefetchOrder = PrefetchOrder.FORWARD_PASS_ORDER,
                          suggested_AG_prefetch_size: Optional[int] = None):
        ids = sorted({self.buffer.param_to_param_group[id(p)] for p in params})
        for b in ids:
            self.async_bucket_gather(b)
        if not prefetch or not ids:
            return
        step = 1 if prefetch_order == PrefetchOrder.FORWARD_PASS_ORDER else -1
        nxt = (max(ids) + 1) if step == 1 else (min(ids) - 1)
        budget = suggested_AG_prefetch_size if suggested_AG_prefetch_size is not None else 0
        inflight = 0
        while 0 <= nxt < self.num_buckets:
            w = self.get_fsdp_buffer(nxt)
            if self.status[nxt] == BucketStatus.EMPTY and w.is_data_distributed:
                if inflight > 0 and inflight + w.bucket_index.size > budget:
                    break
                if not w.temporary_bucket_allocator.can_allocate(nxt):
                    break                                   # double-buffer pool exhausted: this unit is gathered when its turn comes
                self.async_bucket_gather(nxt)
                inflight += w.bucket_index.size
                if inflight >= budget:
                    break
            nxt += step

    def wait_bucket_ready(self, bucket_id: int, empty_ok: bool = False):
        st = self.status[bucket_id]
        if st == BucketStatus.EMPTY:
            if empty_ok:
                return
            raise RuntimeError(f"bucket {bucket_id} was never gathered")
        if st == BucketStatus.COMMUNICATING:
            h = self.handles.pop(bucket_id, None)
            if h is not None:
                h.wait()
            self.status[bucket_id] = BucketStatus.READY_TO_USE
        w = self.get_fsdp_buffer(bucket_id)
        if w.is_data_distributed:
            bucket = w.fetch_bucket()
            for i, p in enumerate(w.params):
                p.data = w.get_item_from_bucket(bucket, i).view(p._fsdp_shape)

    def release_bucket(self, bucket_id: int):
        w = self.get_fsdp_buffer(bucket_id)
        if self.status[bucket_id] == BucketStatus.COMMUNICATING:
            h = self.handles.pop(bucket_id, None)
            if h is not None:
                h.wait()
        if w.is_data_distributed and self.status[bucket_id] != BucketStatus.EMPTY:
            for p in w.params:
                p.data = torch.empty(0, dtype=p.dtype, device=w.device)
            w.free_bucket_storage()
        self.status[bucket_id] = BucketStatus.EMPTY

    def recycle_unused_buckets(self):
        self.reset()


class GradReducePipeline:
    """Per-bucket gradient reduction as soon as a bucket is complete (reference :3820).

    ``reduce_gradients(params)`` marks the parameters' gradients as produced (copying them into the group's full gradient
    bucket, in fp32 when ``grad_reduce_in_fp32``); a complete bucket is reduce-scattered into the main-grad shard (ZeRO-2/3) —
    resident strategies only accumulate and are reduced once at the end of the step."""

    def __init__(self, param_and_grad_buffer: ParamAndGradBuffer, check_nans: bool = False, suggested_queue_capacity: Optional[int] = None):
        self.buffer, self.check_nans = param_and_grad_buffer, check_nans
        self.suggested_queue_capacity = suggested_queue_capacity
        self.ready: Dict[int, set] = defaultdict(set)
        self.inflight: List[Tuple[int, object, torch.Tensor, DataParallelBuffer, int]] = []
        self.reduced_log: List[int] = []

    @property
    def num_buckets(self) -> int:
        return self.buffer.num_buckets

    def get_fsdp_buffer(self, bucket_id: int) -> DataParallelBuffer:
        return self.buffer.parameter_groups[bucket_id].main_grad_buffer

    def reset(self):
        self.wait_for_previous_grad_reduce(0)
        self.ready.clear()

    def wait_for_previous_grad_reduce(self, suggested_queue_size: int = 0):
        """Retire reductions (oldest first) until at most ``suggested_queue_size`` elements remain in flight."""
        def pending():
            return sum(n for *_, n in self.inflight)
        while self.inflight and pending() > suggested_queue_size:
            bid, h, out, gb, _ = self.inflight.pop(0)
            if h is not None:
                h.wait()
            gb.data.add_(out)                                  # fp32 accumulation across micro-batches
            gb.free_bucket_storage()

    @torch.no_grad()
    def reduce_gradients(self, params: Sequence[torch.nn.Parameter], suggested_queue_capacity: Optional[int] = None, async_op: bool = True, force: bool = False):
        cap = suggested_queue_capacity if suggested_queue_capacity is not None else self.suggested_queue_capacity
        touched = set()
        for p in params:
            gi = self.buffer.param_to_param_group[id(p)]
            g = self.buffer.parameter_groups[gi]
            gb = g.main_grad_buffer
            if gb is None or p.grad is None:
                continue
            i = gb.param_idx[id(p)]
            if gb.is_data_distributed:
                bucket = gb.fetch_bucket()
                if not self.ready[gi]:
                    bucket.data.zero_()
                gb.get_item_from_bucket(bucket, i).copy_(p.grad.reshape(-1))
            else:
                gb.get_item(i).add_(p.grad.to(gb.dtype))       # resident full buffer: local accumulation
            p.grad = None
            self.ready[gi].add(i)
            touched.add(gi)
        for gi in sorted(touched):
            g = self.buffer.parameter_groups[gi]
            gb = g.main_grad_buffer
            n_trainable = sum(1 for _ in g.params)
            if (len(self.ready[gi]) < n_trainable and not force) or not gb.is_data_distributed:
                continue
            bucket = gb.fetch_bucket()
            if self.check_nans:
                _check_nan_in_grad(bucket.data, str(gi))
            op = gradient_reduce_preprocessing(bucket.data, gb.gradient_scaling_factor, self.buffer.ddp_config)
            out = torch.empty(gb.data_size, dtype=gb.dtype, device=gb.device)
            if gb.dp_world_size > 1:
                h = dist.reduce_scatter_tensor(out, bucket.data, op=op, group=gb.data_parallel_group, async_op=async_op)
            else:
                out.copy_(bucket.data)
                h = None
            self.inflight.append((gi, h, out, gb, gb.bucket_index.size))
            self.reduced_log.append(gi)
            self.ready[gi] = set()
            if cap is not None:
                self.wait_for_previous_grad_reduce(cap)
