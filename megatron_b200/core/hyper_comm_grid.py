"""n-D named rank grid → process groups (reference ``hyper_comm_grid.py:46-447``).

``HyperCommGrid([2, 2, 2], ["tp", "dp", "pp"])`` lays ranks out with the FIRST dim fastest (rank = offset + sum_i coord_i * stride_i, stride of dim i =
product of the sizes before it); ``create_pg("tp")`` / ``create_pg(["tp", "dp"])`` build (and cache) the groups obtained by varying those dims.

Views: the same rank span can be factorised a second way (``register_view("expert", [2, 4], ["etp", "ep"], shared_dims=[...])`` — e.g. the dense
tp×cp×dp layout next to the expert etp×ep×edp layout of a MoE model).  A dim listed in ``shared_dims`` must produce the SAME rank groups in both
factorisations; its group is then created once and shared.  Everything that enumerates ranks is a pure function of (shape, names, offset) and is
unit-tested without a process group."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch.distributed as dist

BASE_VIEW = "__base__"


@dataclass(frozen=True)
class _View:
    name: str
    shape: Tuple[int, ...]
    dim_names: Tuple[str, ...]
    shared_dims: Tuple[str, ...] = field(default_factory=tuple)

    def strides(self) -> Dict[str, int]:
        out, s = {}, 1
        for n, k in zip(self.dim_names, self.shape):
            out[n] = s
            s *= k
        return out

    def canonical(self, dims: Union[str, Sequence[str]]) -> List[str]:
        """Requested dims, slowest-varying first (the order the group key is spelled in: ``"dp-tp"`` for dim_names ``[tp, .., dp]``)."""
        dims = [dims] if isinstance(dims, str) else list(dims)
        if len(set(dims)) != len(dims):
            raise ValueError(f"duplicate dims in {dims}")
        for d in dims:
            if d not in self.dim_names:
                raise ValueError(f"{d!r} is not in view {self.name!r} with dim_names {list(self.dim_names)}")
        return sorted(dims, key=self.dim_names.index, reverse=True)


def enumerate_groups(shape: Sequence[int], dim_names: Sequence[str], dims: Sequence[str], rank_offset: int = 0) -> List[List[int]]:
    """Rank groups obtained by varying ``dims`` (members in ascending rank order, groups ordered by their first member)."""
    view = _View("_", tuple(shape), tuple(dim_names))
    stride, size = view.strides(), dict(zip(dim_names, shape))
    inside = sorted(dims, key=list(dim_names).index)            # fastest first
    outside = [d for d in dim_names if d not in dims]

    def offsets(names):                                           # all sum_i c_i * stride_i over the given dims, first name fastest
        out = [0]
        for n in names:
            out = [o + c * stride[n] for c in range(size[n]) for o in out]
        return out

    member_off = sorted(offsets(inside))
    return [[rank_offset + base + m for m in member_off] for base in sorted(offsets(outside))]


class HyperCommGrid:
    def __init__(self, shape: Sequence[int], dim_names: Sequence[str], rank_offset: int = 0, backend: Optional[str] = None):
        if len(shape) != len(dim_names):
            raise ValueError(f"len(shape) {shape} != len(dim_names) {dim_names}")
        if len(set(dim_names)) != len(dim_names):
            raise ValueError("dimension names must be unique")
        if rank_offset < 0:
            raise ValueError(f"rank_offset must be non-negative, got {rank_offset}")
        self.shape, self.dim_names, self.rank_offset, self.backend = list(shape), list(dim_names), rank_offset, backend
        self.size = 1
        for s in shape:
            self.size *= int(s)
        world = self._world_size()
        if world is not None and rank_offset + self.size > world:
            raise RuntimeError(f"grid of {self.size} ranks at offset {rank_offset} exceeds world size {world}")
        self._views: Dict[str, _View] = {BASE_VIEW: _View(BASE_VIEW, tuple(self.shape), tuple(self.dim_names))}
        # base groups (and shared dims of any view) are keyed "dp-tp"; view-private groups (view name, "ep-etp")
        self._pgs: Dict[Union[str, Tuple[str, str]], Optional[dist.ProcessGroup]] = {}

    @staticmethod
    def _world_size() -> Optional[int]:
        if "WORLD_SIZE" in os.environ:
            return int(os.environ["WORLD_SIZE"])
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return None

    # ---- views --------------------------------------------------------------------------------------------------------------------
    def register_view(self, name: str, shape: Sequence[int], dim_names: Sequence[str], shared_dims: Optional[Sequence[str]] = None) -> None:
        if name in self._views:
            raise ValueError(f"view {name!r} is already registered")
        if len(shape) != len(dim_names):
            raise ValueError(f"len(shape) {shape} != len(dim_names) {dim_names}")
        if len(set(dim_names)) != len(dim_names):
            raise ValueError(f"view {name!r} has duplicate dim_names: {dim_names}")
        if any((not isinstance(s, int)) or isinstance(s, bool) or s <= 0 for s in shape):
            raise ValueError(f"view {name!r} shape must be positive ints, got {shape}")
        n = 1
        for s in shape:
            n *= s
        if n != self.size:
            raise ValueError(f"view {name!r} shape {list(shape)} has size {n}, but the grid size is {self.size}")
        shared = list(shared_dims or [])
        if len(set(shared)) != len(shared):
            raise ValueError(f"view {name!r} has duplicate shared_dims: {shared}")
        for d in shared:
            if d not in self.dim_names:
                raise ValueError(f"shared dim {d!r} of view {name!r} is not in the base view {self.dim_names}")
            if d not in dim_names:
                raise ValueError(f"shared dim {d!r} of view {name!r} is not in the view's dim_names {list(dim_names)}")
        # each shared dim alone, and all of them together, must enumerate to the same groups under both factorisations
        for probe in [[d] for d in shared] + ([shared] if len(shared) > 1 else []):
            a = enumerate_groups(self.shape, self.dim_names, probe, self.rank_offset)
            b = enumerate_groups(shape, dim_names, probe, self.rank_offset)
            if a != b:
                raise ValueError(f"shared dims {probe} have different membership across views: base {a} != view {name!r} {b}")
        self._views[name] = _View(name, tuple(shape), tuple(dim_names), tuple(shared))

    def _view(self, view: Optional[str]) -> _View:
        key = BASE_VIEW if view is None else view
        if key not in self._views:
            raise KeyError(f"view {key!r} is not registered; registered: {sorted(self._views)}")
        return self._views[key]

    def _pg_key(self, view: _View, canon: List[str]):
        """Groups over shared dims only live under the base key (created once, reachable from both views)."""
        if view.name == BASE_VIEW or all(d in view.shared_dims for d in canon):
            base = self._views[BASE_VIEW]
            return "-".join(base.canonical(canon)), base
        return (view.name, "-".join(canon)), view

    # ---- enumeration --------------------------------------------------------------------------------------------------------------
    def get_rank_enum(self, dims: Union[str, Sequence[str]], *, view: Optional[str] = None) -> List[List[int]]:
        v = self._view(view)
        return enumerate_groups(v.shape, v.dim_names, v.canonical(dims), self.rank_offset)

    def is_current_rank_in_grid(self) -> bool:
        rank = dist.get_rank()
        return self.rank_offset <= rank < self.rank_offset + self.size

    def coords(self, rank: int, *, view: Optional[str] = None) -> Dict[str, int]:
        """Coordinate of a global rank in the grid (ours; handy for 'am I the first/last stage' style questions)."""
        v, r = self._view(view), rank - self.rank_offset
        if not 0 <= r < self.size:
            raise ValueError(f"rank {rank} is outside the grid [{self.rank_offset}, {self.rank_offset + self.size})")
        out = {}
        for n, k in zip(v.dim_names, v.shape):
            out[n], r = r % k, r // k
        return out

    # ---- process groups ------------------------------------------------------------------------------------------------------------
    def create_pg(self, dims: Union[str, Sequence[str]], *, view: Optional[str] = None, **kwargs) -> Optional[dist.ProcessGroup]:
        v = self._view(view)
        key, enum_view = self._pg_key(v, v.canonical(dims))
        if key in self._pgs:
            raise KeyError(f"process group {dims} (view {v.name!r}) has already been created; options cannot be compared, use get_pg")
        enum = enumerate_groups(enum_view.shape, enum_view.dim_names, enum_view.canonical(dims), self.rank_offset)
        mine, rank = None, dist.get_rank()
        for ranks in enum:                                         # every rank of the world walks every group (new_group is collective)
            pg = dist.new_group(ranks, backend=self.backend, **kwargs)
            if rank in ranks:
                mine = pg
        self._pgs[key] = mine
        return mine

    def get_pg(self, dims: Union[str, Sequence[str]], *, view: Optional[str] = None) -> dist.ProcessGroup:
        v = self._view(view)
        key, _ = self._pg_key(v, v.canonical(dims))
        if key not in self._pgs:
            raise KeyError(f"process group for {key} has not been created; call create_pg first")
        return self._pgs[key]

    def destroy(self) -> None:
        """Tear down every group this grid created (a group shared by two views is stored once, so it is destroyed once)."""
        seen = set()
        for pg in self._pgs.values():
            if pg is not None and id(pg) not in seen:
                seen.add(id(pg))
                try:
                    dist.destroy_process_group(pg)
                except (ValueError, RuntimeError, AssertionError):   # already gone with the default group
                    pass
        self._pgs.clear()
