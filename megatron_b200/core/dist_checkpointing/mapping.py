"""Sharded state-dict leaves (reference ``dist_checkpointing/mapping.py:48-470``).

A *sharded state dict* is a nested dict/list whose leaves describe how a local
tensor fits into a global, parallelism-independent tensor: ``ShardedTensor``
(key, local data, global shape, global offset, fragmentation per axis,
replica id).  ``replica_id`` all-zero marks the single writer.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field, replace
from itertools import chain
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch

ReplicaId = Union[int, Tuple[int, ...]]
StateDict = Dict[str, Any]
ShardedStateDict = Dict[str, Any]


class ShardedBase(ABC):
    key: str
    data: object
    replica_id: ReplicaId

    @abstractmethod
    def validate_metadata_integrity(self):
        ...

    @abstractmethod
    def without_data(self) -> "ShardedBase":
        ...


@dataclass
class ShardedTensor(ShardedBase):
    key: str
    data: Optional[torch.Tensor] = field(repr=False)
    dtype: torch.dtype = None
    local_shape: Tuple[int, ...] = ()
    global_shape: Tuple[int, ...] = ()
    global_offset: Tuple[int, ...] = ()
    axis_fragmentations: Optional[Tuple[int, ...]] = None
    replica_id: ReplicaId = 0
    prepend_axis_num: int = 0
    allow_shape_mismatch: bool = False
    flattened_range: Optional[slice] = None

    def __post_init__(self):
        self.validate_metadata_integrity()

    def validate_metadata_integrity(self):
        has_flat = self.flattened_range is not None
        if self.data is not None:
            if self.data.dtype != self.dtype:
                raise ValueError(f"data dtype {self.data.dtype} != declared dtype {self.dtype} for {self.key}")
            if not has_flat and tuple(self.data.shape) != tuple(self.local_shape):
                raise ValueError(f"data shape {tuple(self.data.shape)} != local_shape {self.local_shape} for {self.key}")
        if len(self.global_shape) != len(self.global_offset):
            raise ValueError(f"global offset rank must match global shape rank for {self.key}")
        if len(self.local_shape) + self.prepend_axis_num != len(self.global_shape):
            raise ValueError(f"local rank + prepend_axis_num must equal global rank for {self.key}")
        for off, sh in zip(self.global_offset[self.prepend_axis_num :], self.local_shape):
            if sh and off % sh != 0 and not self.allow_shape_mismatch:
                raise ValueError(f"global offset {self.global_offset} must be a multiple of local shape {self.local_shape} for {self.key}")

    def global_slice(self) -> Tuple[Union[int, slice], ...]:
        return tuple(
            chain(
                (off for off in self.global_offset[: self.prepend_axis_num]),
                (slice(off, off + sh) for off, sh in zip(self.global_offset[self.prepend_axis_num :], self.local_shape)),
            )
        )

    def local_chunk_offset_in_global(self) -> Tuple[int, ...]:
        out = []
        for i, off in enumerate(self.global_offset):
            if i < self.prepend_axis_num:
                out.append(off)
            else:
                sh = self.local_shape[i - self.prepend_axis_num]
                out.append(off // sh if sh else 0)
        return tuple(out)

    def max_allowed_chunks(self) -> Tuple[int, ...]:
        out = []
        for i, g in enumerate(self.global_shape):
            if i < self.prepend_axis_num:
                out.append(g)
            else:
                sh = self.local_shape[i - self.prepend_axis_num]
                out.append(g // sh if sh else 1)
        return tuple(out)

    def without_data(self):
        return replace(self, data=None)

    @classmethod
    def from_rank_offsets(
        cls,
        key: str,
        data: torch.Tensor,
        *rank_offsets: Tuple[int, int, int],
        replica_id: ReplicaId = 0,
        prepend_axis_num: int = 0,
        flattened_range: None = None,
        **init_kwargs,
    ):
        """Each ``(axis, rank_offset, axis_fragm)`` says: along ``axis`` the global tensor is
        cut into ``axis_fragm`` equal pieces and this shard is piece ``rank_offset``."""
        if flattened_range is not None:
            raise ValueError("flattened_range is not supported in from_rank_offsets")
        ndim = data.ndim + prepend_axis_num
        global_offset = [0] * ndim
        global_shape = ([1] * prepend_axis_num) + list(data.shape)
        fragm = [1] * ndim
        seen = set()
        for axis, off, n in rank_offsets:
            if axis < 0 or off < 0 or n < 1 or off >= n:
                raise ValueError(f"invalid rank offset ({axis}, {off}, {n}) for {key}")
            if axis in seen:
                raise ValueError(f"duplicate axis {axis} in rank offsets for {key}")
            seen.add(axis)
            local = 1 if axis < prepend_axis_num else data.shape[axis - prepend_axis_num]
            global_shape[axis] = n * local
            global_offset[axis] = off * local
            fragm[axis] = n
        return cls(
            key, data, data.dtype, tuple(data.shape), tuple(global_shape), tuple(global_offset), tuple(fragm),
            replica_id, prepend_axis_num, **init_kwargs,
        )

    def init_data(self, device, init_fn=torch.empty):
        if self.data is None:
            self.data = init_fn(self.local_shape, dtype=self.dtype, device=device)

    def narrow(self, dim: int, start: int, length: int):
        """Sub-shard along a local dim (used by the fully-parallel save to split work)."""
        gdim = dim + self.prepend_axis_num
        go = list(self.global_offset)
        go[gdim] += start
        ls = list(self.local_shape)
        ls[dim] = length
        frag = list(self.axis_fragmentations) if self.axis_fragmentations else None
        data = self.data.narrow(dim, start, length) if self.data is not None else None
        return replace(self, data=data, local_shape=tuple(ls), global_offset=tuple(go), axis_fragmentations=tuple(frag) if frag else None, allow_shape_mismatch=True)


def is_main_replica(replica_id: ReplicaId) -> bool:
    if isinstance(replica_id, int):
        return replica_id == 0
    return all(r == 0 for r in replica_id)


class LocalNonpersistentObject:
    """Kept out of the checkpoint; returned as-is on load."""

    def __init__(self, obj):
        self.obj = obj

    def unwrap(self):
        return self.obj


@dataclass
class ShardedObject(ShardedBase):
    """Arbitrary picklable object addressed by (key, global_offset) — e.g. RNG state per rank."""

    key: str
    data: object
    global_shape: Tuple[int, ...]
    global_offset: Tuple[int, ...]
    replica_id: ReplicaId = 0

    def __post_init__(self):
        self.validate_metadata_integrity()

    def validate_metadata_integrity(self):
        if len(self.global_shape) != len(self.global_offset):
            raise ValueError(f"global offset rank must match global shape rank for {self.key}")

    def without_data(self):
        return replace(self, data=None)

    @property
    def unique_key(self):
        return f"{self.key}/shard_{'.'.join(map(str, self.global_offset))}_{'.'.join(map(str, self.global_shape))}"

    def __str__(self):
        return f"{type(self).__name__}(key='{self.key}')"

    @classmethod
    def empty_from_unique_key(cls, unique_key, replica_id: ReplicaId = 0):
        key, shard = unique_key.split("/")
        _, off, shp = shard.split("_")
        return cls(key, None, tuple(map(int, shp.split("."))), tuple(map(int, off.split("."))), replica_id)


FactoryBuildFn = Callable[[str, torch.Tensor, ReplicaId, Optional[slice]], ShardedStateDict]
FactoryMergeFn = Callable[[StateDict], torch.Tensor]


@dataclass
class ShardedTensorFactory(ShardedBase):
    """Deferred transformation: ``build_fn`` expands one in-memory tensor into several
    ShardedTensors at save/load time (e.g. SwiGLU fc1 → two TP-sharded halves) and
    ``merge_fn`` folds the loaded pieces back."""

    key: str
    data: torch.Tensor
    build_fn: FactoryBuildFn
    merge_fn: FactoryMergeFn
    replica_id: ReplicaId = 0
    flattened_range: Optional[slice] = None

    def build(self):
        return self.build_fn(self.key, self.data, self.replica_id, self.flattened_range)

    def validate_metadata_integrity(self):
        pass

    def without_data(self):
        return replace(self, data=None)


def apply_factories(sharded_state_dict: ShardedStateDict):
    """Expand every ``ShardedTensorFactory`` in place."""

    def rec(x):
        if isinstance(x, ShardedTensorFactory):
            return rec(x.build())
        if isinstance(x, dict):
            for k in list(x.keys()):
                x[k] = rec(x[k])
        elif isinstance(x, list):
            for i in range(len(x)):
                x[i] = rec(x[i])
        return x

    rec(sharded_state_dict)


def apply_factory_merges(x1: StateDict, x2: ShardedStateDict, key: Tuple[str, ...] = ()):
    """Walk the loaded state dict ``x1`` alongside the pre-factory template ``x2`` and merge."""
    if isinstance(x2, ShardedTensorFactory):
        return x2.merge_fn(x1)
    if isinstance(x1, dict) and isinstance(x2, dict):
        for k, v2 in x2.items():
            if k not in x1:
                raise ValueError(f"different dict keys encountered in apply_factory_merges at {key}: missing {k}")
            x1[k] = apply_factory_merges(x1[k], v2, key + (k,))
    elif isinstance(x1, list) and isinstance(x2, list):
        if len(x1) != len(x2):
            raise ValueError(f"list length mismatch in apply_factory_merges at {key}")
        for i, v2 in enumerate(x2):
            x1[i] = apply_factory_merges(x1[i], v2, key + (i,))
    elif isinstance(x1, list) and isinstance(x2, dict):
        for k, v2 in x2.items():
            x1[k] = apply_factory_merges(x1[k], v2, key + (k,))
    return x1
