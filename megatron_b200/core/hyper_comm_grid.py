"""n-D named rank grid → process groups (reference ``hyper_comm_grid.py:46-447``).

``HyperCommGrid([2, 2, 2], ["tp", "dp", "pp"])`` lays ranks out with the FIRST dim fastest;
``create_pg("tp")`` / ``create_pg(["tp", "dp"])`` build (and cache) the groups obtained by
varying those dims; ``get_rank_enum`` is the pure-function part and is unit-tested on CPU."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch.distributed as dist


class HyperCommGrid:
    def __init__(self, shape: Sequence[int], dim_names: Sequence[str], rank_offset: int = 0, backend: Optional[str] = None):
        if len(shape) != len(dim_names):
            raise ValueError(f"len(shape) {shape} != len(dim_names) {dim_names}")
        if len(set(dim_names)) != len(dim_names):
            raise ValueError("dimension names must be unique")
        self.shape, self.dim_names, self.rank_offset, self.backend = list(shape), list(dim_names), rank_offset, backend
        self.size = int(np.prod(shape))
        if dist.is_available() and dist.is_initialized() and rank_offset + self.size > dist.get_world_size():
            raise RuntimeError(f"grid of {self.size} ranks at offset {rank_offset} exceeds world size {dist.get_world_size()}")
        self._pgs: Dict[str, dist.ProcessGroup] = {}

    def _key(self, dims: Union[str, Sequence[str]]) -> List[str]:
        dims = [dims] if isinstance(dims, str) else list(dims)
        for d in dims:
            if d not in self.dim_names:
                raise KeyError(f"unknown dim {d}; have {self.dim_names}")
        return sorted(dims, key=self.dim_names.index)

    def get_rank_enum(self, dims: Union[str, Sequence[str]]) -> List[List[int]]:
        dims = self._key(dims)
        n = len(self.shape)
        grid = np.arange(self.size).reshape(list(reversed(self.shape))) + self.rank_offset
        ax = lambda name: n - 1 - self.dim_names.index(name)  # noqa: E731
        masked = sorted(ax(d) for d in dims)
        rest = sorted(a for a in range(n) if a not in masked)
        gsize = int(np.prod([grid.shape[a] for a in masked]))
        out = grid.transpose(rest + masked).reshape(-1, gsize)
        groups = [list(map(int, r)) for r in out]
        groups.sort(key=lambda g: g[0])
        return groups

    def create_pg(self, dims: Union[str, Sequence[str]], **kwargs) -> Optional[dist.ProcessGroup]:
        key = "-".join(self._key(dims))
        if key in self._pgs:
            raise KeyError(f"process group for {key} already exists; use get_pg")
        mine = None
        rank = dist.get_rank()
        for ranks in self.get_rank_enum(dims):
            pg = dist.new_group(ranks, backend=self.backend, **kwargs)
            if rank in ranks:
                mine = pg
        self._pgs[key] = mine
        return mine

    def get_pg(self, dims: Union[str, Sequence[str]]) -> dist.ProcessGroup:
        key = "-".join(self._key(dims))
        if key not in self._pgs:
            raise KeyError(f"process group for {key} has not been created; call create_pg first")
        return self._pgs[key]
