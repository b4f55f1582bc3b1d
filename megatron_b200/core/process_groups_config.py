"""Explicit process-group bundle (reference ``process_groups_config.py:27-202``).

Every module accepts ``pg_collection=``; ``ProcessGroupCollection.use_mpu_process_groups()``
builds one from the global registry in ``parallel_state``."""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import List, Optional

import torch.distributed as dist

from . import parallel_state as ps

_FIELD_TO_REGISTRY = {
    "tp": "tp", "pp": "pp", "cp": "cp", "dp": "dp", "dp_cp": "dp_cp", "mp": "mp", "embd": "embd", "pos_embd": "pos_embd", "tp_cp": "tp_cp",
    "tp_dp": "tp_dp", "tp_dp_cp": "tp_dp_cp", "ep": "ep", "expt_tp": "expt_tp", "tp_ep": "tp_ep", "tp_ep_pp": "tp_ep_pp", "expt_dp": "expt_dp",
    "intra_dp_cp": "intra_dp_cp", "inter_dist_opt": "inter_dist_opt",
}


@dataclass
class ProcessGroupCollection:
    tp: Optional[dist.ProcessGroup] = field(default=None)
    pp: Optional[dist.ProcessGroup] = field(default=None)
    cp: Optional[dist.ProcessGroup] = field(default=None)
    dp: Optional[dist.ProcessGroup] = field(default=None)
    dp_cp: Optional[dist.ProcessGroup] = field(default=None)
    mp: Optional[dist.ProcessGroup] = field(default=None)
    embd: Optional[dist.ProcessGroup] = field(default=None)
    pos_embd: Optional[dist.ProcessGroup] = field(default=None)
    tp_cp: Optional[dist.ProcessGroup] = field(default=None)
    tp_dp: Optional[dist.ProcessGroup] = field(default=None)
    tp_dp_cp: Optional[dist.ProcessGroup] = field(default=None)
    ep: Optional[dist.ProcessGroup] = field(default=None)
    expt_tp: Optional[dist.ProcessGroup] = field(default=None)
    tp_ep: Optional[dist.ProcessGroup] = field(default=None)
    tp_ep_pp: Optional[dist.ProcessGroup] = field(default=None)
    expt_dp: Optional[dist.ProcessGroup] = field(default=None)
    intra_dp_cp: Optional[dist.ProcessGroup] = field(default=None)
    inter_dist_opt: Optional[dist.ProcessGroup] = field(default=None)
    hcp: Optional[List[dist.ProcessGroup]] = field(default=None)

    @classmethod
    def use_mpu_process_groups(cls, required_pgs: Optional[List[str]] = None) -> "ProcessGroupCollection":
        names = required_pgs or list(_FIELD_TO_REGISTRY)
        unknown = [n for n in names if n not in _FIELD_TO_REGISTRY and n != "hcp"]
        if unknown:
            raise ValueError(f"invalid process groups requested: {unknown}")
        kw = {n: ps.get_group(_FIELD_TO_REGISTRY[n], check_initialized=False) for n in names if n != "hcp"}
        if "hcp" in names:
            kw["hcp"] = ps.get_hierarchical_context_parallel_groups(check_initialized=False)
        return cls(**kw)

    def __repr__(self):
        active = [f.name for f in fields(self) if getattr(self, f.name) is not None]
        return f"ProcessGroupCollection({', '.join(active)})"


@dataclass
class MultiModuleProcessGroupCollection:
    """One collection per sub-model for multi-module (MIMO) pipelines on different grids (reference :718)."""

    module_pgs: dict = field(default_factory=dict)
    language_model_module_name: Optional[str] = None

    def get_language_model_cp_size(self) -> int:
        pg = self.module_pgs.get(self.language_model_module_name)
        return dist.get_world_size(pg.cp) if pg is not None and pg.cp is not None else 1
