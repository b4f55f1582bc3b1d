"""Explicit process-group bundle (reference ``process_groups_config.py:27-202``).

Every module accepts ``pg_collection=``; ``ProcessGroupCollection.use_mpu_process_groups()``
builds one from the global registry in ``parallel_state``."""
from __future__ import annotations

from dataclasses import dataclass, field, fields
import logging
from typing import Dict, Iterator, List, Optional

import torch.distributed as dist

from . import parallel_state as ps

logger = logging.getLogger(__name__)

_FIELD_TO_REGISTRY = {
    "tp": "tp", "pp": "pp", "cp": "cp", "dp": "dp", "dp_cp": "dp_cp", "mp": "mp", "embd": "embd", "pos_embd": "pos_embd", "tp_cp": "tp_cp",
    "tp_dp": "tp_dp", "tp_dp_cp": "tp_dp_cp", "ep": "ep", "expt_tp": "expt_tp", "tp_ep": "tp_ep", "tp_ep_pp": "tp_ep_pp", "expt_dp": "expt_dp",
    "intra_dp_cp": "intra_dp_cp", "inter_dist_opt": "inter_dist_opt", "intra_expt_dp": "expt_dp", "intra_dist_opt": "intra_dp_cp", "expt_tp_pp": "tp_ep_pp",
    "gtp_remat": "gtp_remat", "expt_gtp_remat": "egtp_remat",
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
    intra_expt_dp: Optional[dist.ProcessGroup] = field(default=None)
    intra_dist_opt: Optional[dist.ProcessGroup] = field(default=None)
    expt_tp_pp: Optional[dist.ProcessGroup] = field(default=None)
    gtp_remat: Optional[dist.ProcessGroup] = field(default=None)
    expt_gtp_remat: Optional[dist.ProcessGroup] = field(default=None)
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
        parts = []
        for f in fields(self):
            pg = getattr(self, f.name)
            if pg is None:
                continue
            parts.append(f"{f.name}({[g.size() for g in pg]})" if isinstance(pg, list) else f"{f.name}({pg.size()})")
        return f"ProcessGroupCollection({', '.join(parts)})" if parts else "ProcessGroupCollection(empty)"

    # ---- from a HyperCommGrid ------------------------------------------------------------------------------------------------------
    @classmethod
    def from_grid(cls, grid, expert_view: Optional[str] = None, create: bool = True) -> "ProcessGroupCollection":
        """Build the collection a model on ``grid`` needs (ours; the reference leaves this to the caller, e.g. examples/mimo).  Dense groups come
        from the base view dims among {tp, cp, dp, pp}; with ``expert_view`` (a registered view with dims among {expt_tp, ep, expt_dp, pp}) the
        expert groups as well.  Missing dims give ``None`` (world-size-1 semantics in the consumers)."""
        have = set(grid.dim_names)

        def pg(dims, view=None):
            names = set(grid.dim_names if view is None else grid._view(view).dim_names)
            dims = [d for d in dims if d in names]
            if not dims:
                return None
            try:
                return grid.get_pg(dims, view=view)
            except KeyError:
                if not create:
                    raise
                return grid.create_pg(dims, view=view)

        kw = dict(tp=pg(["tp"]), cp=pg(["cp"]), dp=pg(["dp"]), pp=pg(["pp"]))
        kw["dp_cp"] = pg(["dp", "cp"]) if "cp" in have else kw["dp"]
        kw["tp_cp"] = pg(["tp", "cp"]) if "cp" in have else kw["tp"]
        kw["mp"] = pg(["tp", "pp"]) if "pp" in have else kw["tp"]
        kw["tp_dp_cp"] = pg(["tp", "dp", "cp"])
        kw["intra_dp_cp"] = kw["dp_cp"]
        if expert_view is not None:
            kw.update(ep=pg(["ep"], expert_view), expt_tp=pg(["expt_tp"], expert_view), expt_dp=pg(["expt_dp"], expert_view),
                      tp_ep=pg(["expt_tp", "ep"], expert_view), tp_ep_pp=pg(["expt_tp", "ep", "pp"], expert_view),
                      expt_tp_pp=pg(["expt_tp", "pp"], expert_view))
            kw["intra_expt_dp"] = kw["expt_dp"]
        return cls(**kw)

    # ---- group bundles for the optimizer / DDP (reference :337-690) ----------------------------------------------------------------
    @staticmethod
    def is_gtp_remat_active(groups: Dict) -> bool:
        return any(groups.get(k) is not None and groups[k].size() > 1 for k in ("gtp_remat_group", "expt_gtp_remat_group"))

    @staticmethod
    def _from_parallel_state(num_instances: int, use_distributed_optimizer: bool, gloo: bool) -> Dict:
        g = lambda name: ps.get_group(name, check_initialized=False)  # noqa: E731
        gtp, egtp = ps.get_gtp_weight_remat_group(check_initialized=False), ps.get_expert_gtp_weight_remat_group(check_initialized=False)
        out = dict(dp_group=g("dp"), dp_cp_group=g("dp_cp"), intra_dp_cp_group=g("intra_dp_cp"), expt_dp_group=g("expt_dp"), intra_expt_dp_group=g("expt_dp"),
                   mp_group=g("mp"), expt_tp_pp_group=g("tp_ep_pp"), tp_group=g("tp"), pp_group=g("pp"), ep_group=g("ep"), gtp_remat_group=gtp,
                   expt_gtp_remat_group=egtp, inter_dist_opt_group=g("inter_dist_opt") if num_instances > 1 else None,
                   intra_dist_opt_group=g("intra_dp_cp") if use_distributed_optimizer or num_instances > 1 else None)
        gtp_on = ProcessGroupCollection.is_gtp_remat_active(out)
        out["intra_dp_cp_group_gloo"] = ps.get_data_parallel_group_gloo(with_context_parallel=True, partial_data_parallel=True) if gloo and not gtp_on else None
        out["intra_expt_dp_group_gloo"] = ps.get_expert_data_parallel_group_gloo(partial_expert_data_parallel=True) if gloo and not gtp_on else None
        return out

    def _resolve_data_groups(self, cp_size: int, num_instances: int, create_missing_expt_dp: bool) -> Dict:
        """The fallbacks shared by the two public helpers: dp required; dp_cp := dp when cp == 1; one optimizer instance → the intra groups ARE the
        full ones; several instances → intra_dp_cp / intra_expt_dp / inter_dist_opt must be given."""
        if self.dp is None:
            raise ValueError("dp process group is required but not provided in pg_collection")
        out = {"dp_group": self.dp}
        if self.dp_cp is not None:
            out["dp_cp_group"] = self.dp_cp
        elif cp_size == 1:
            out["dp_cp_group"] = self.dp
        else:
            raise ValueError("dp_cp process group is required when context_parallel_size > 1 but not provided in pg_collection")
        if self.expt_dp is not None:
            out["expt_dp_group"] = self.expt_dp
        elif create_missing_expt_dp:
            logger.warning("no expert data parallel group in pg_collection: using a group of just this rank")
            out["expt_dp_group"] = dist.new_group(ranks=[dist.get_rank()], use_local_synchronization=True)
        else:
            raise ValueError("expt_dp process group is required but not provided in pg_collection")
        if num_instances == 1:
            out.update(intra_dp_cp_group=out["dp_cp_group"], intra_expt_dp_group=out["expt_dp_group"], inter_dist_opt_group=None)
        else:
            if self.intra_dp_cp is None or self.intra_expt_dp is None or self.inter_dist_opt is None:
                raise ValueError("intra_dp_cp, intra_expt_dp, and inter_dist_opt process groups are required when using multiple optimizer instances "
                                 "(>1) but not provided in pg_collection")
            out.update(intra_dp_cp_group=self.intra_dp_cp, intra_expt_dp_group=self.intra_expt_dp, inter_dist_opt_group=self.inter_dist_opt)
        out["gtp_remat_group"], out["expt_gtp_remat_group"] = self.gtp_remat, self.expt_gtp_remat
        return out

    @staticmethod
    def setup_process_groups_for_optimizer(pg_collection: Optional["ProcessGroupCollection"], model_chunks: List, use_gloo_process_groups: bool = True) -> Dict:
        """Every group the optimizer stack needs, as a dict (``dp_group``, ``dp_cp_group``, ``intra_dp_cp_group``, ``expt_dp_group``,
        ``intra_expt_dp_group``, ``mp_group``, ``expt_tp_pp_group``, ``inter_dist_opt_group``, ``intra_dist_opt_group``, the two ``*_gloo`` groups).
        ``pg_collection=None`` → the global registry."""
        ddp_cfg = getattr(model_chunks[0], "ddp_config", None) if model_chunks else None
        n_inst = getattr(ddp_cfg, "num_distributed_optimizer_instances", 1) if ddp_cfg is not None else 1
        use_do = getattr(ddp_cfg, "use_distributed_optimizer", False) if ddp_cfg is not None else False
        if pg_collection is None:
            out = ProcessGroupCollection._from_parallel_state(n_inst, True, use_gloo_process_groups)
            for k in ("tp_group", "pp_group", "ep_group"):
                out.pop(k)
            return out
        cfg = getattr(model_chunks[0], "config", None) if model_chunks else None
        out = pg_collection._resolve_data_groups(getattr(cfg, "context_parallel_size", 1), n_inst, create_missing_expt_dp=False)
        if pg_collection.mp is None or (pg_collection.expt_tp_pp is None and pg_collection.tp_ep_pp is None):
            raise ValueError("mp and expt_tp_pp process groups are required but not provided in pg_collection")
        out["mp_group"] = pg_collection.mp
        out["expt_tp_pp_group"] = pg_collection.expt_tp_pp if pg_collection.expt_tp_pp is not None else pg_collection.tp_ep_pp
        out["intra_dist_opt_group"] = pg_collection.intra_dist_opt if pg_collection.intra_dist_opt is not None else (out["intra_dp_cp_group"] if use_do or n_inst > 1 else None)
        # gloo twins are not derivable from a user-supplied collection: callers that want CPU-side gathers pass gloo groups themselves
        out["intra_dp_cp_group_gloo"] = out["intra_expt_dp_group_gloo"] = None
        return out

    @staticmethod
    def setup_process_groups_for_ddp(pg_collection: Optional["ProcessGroupCollection"], config, ddp_config) -> Dict:
        """Groups for the DDP wrapper: the data groups plus ``tp_group`` / ``pp_group`` / ``ep_group`` (it needs them to tell expert from dense buffers
        and for the embedding all-reduces)."""
        n_inst = getattr(ddp_config, "num_distributed_optimizer_instances", 1)
        if pg_collection is None:
            out = ProcessGroupCollection._from_parallel_state(n_inst, getattr(ddp_config, "use_distributed_optimizer", False), gloo=False)
            for k in ("mp_group", "expt_tp_pp_group", "intra_dp_cp_group_gloo", "intra_expt_dp_group_gloo"):
                out.pop(k)
            return out
        out = pg_collection._resolve_data_groups(getattr(config, "context_parallel_size", 1), n_inst, create_missing_expt_dp=True)
        if pg_collection.tp is None or pg_collection.pp is None or pg_collection.ep is None:
            raise ValueError("tp, pp and ep process groups are required but not provided in pg_collection")
        out.update(tp_group=pg_collection.tp, pp_group=pg_collection.pp, ep_group=pg_collection.ep)
        return out


@dataclass
class MultiModuleProcessGroupCollection:
    """One collection per sub-model for multi-module (MIMO) pipelines whose modules sit on different grids (reference :718-850).  Dict-like over module
    names; ``language_model_module_name`` marks the LLM (``None`` when this rank hosts no LLM, e.g. an encoder-only rank)."""

    module_pgs: Dict[str, ProcessGroupCollection] = field(default_factory=dict)
    language_model_module_name: Optional[str] = None

    def __post_init__(self):
        if not self.module_pgs:
            raise ValueError("module_pgs dict cannot be empty")
        if self.language_model_module_name is not None and self.language_model_module_name not in self.module_pgs:
            raise ValueError(f"language_model_module_name {self.language_model_module_name!r} not found in module_pgs keys: {list(self.module_pgs)}")

    @classmethod
    def from_grids(cls, module_to_grid: Dict[str, "object"], language_model_module_name: Optional[str] = None, expert_views: Optional[Dict[str, str]] = None):
        """Collections for the modules whose grid contains THIS rank (ours).  Every rank must call it (group creation is collective)."""
        rank, out = dist.get_rank(), {}
        for name, grid in module_to_grid.items():
            pgc = ProcessGroupCollection.from_grid(grid, expert_view=(expert_views or {}).get(name))
            if grid.rank_offset <= rank < grid.rank_offset + grid.size:
                out[name] = pgc
        lm = language_model_module_name if language_model_module_name in out else None
        return cls(out, lm)

    def has_language_model(self) -> bool:
        return self.language_model_module_name is not None

    def get_language_model_collection(self) -> ProcessGroupCollection:
        if self.language_model_module_name is None:
            raise ValueError("No language model specified for this collection")
        return self.module_pgs[self.language_model_module_name]

    def get_language_model_cp_size(self) -> int:
        cp = self.get_language_model_collection().cp
        return cp.size() if cp is not None else 1

    def get_module_collection(self, module_name: str) -> ProcessGroupCollection:
        if module_name not in self.module_pgs:
            raise KeyError(f"module {module_name!r} not found; available: {list(self.module_pgs)}")
        return self.module_pgs[module_name]

    def __len__(self) -> int:
        return len(self.module_pgs)

    def __getitem__(self, module_name: str) -> ProcessGroupCollection:
        return self.get_module_collection(module_name)

    def __iter__(self) -> Iterator[str]:
        return iter(self.module_pgs)

    def __contains__(self, module_name) -> bool:
        return module_name in self.module_pgs

    def keys(self):
        return self.module_pgs.keys()

    def values(self):
        return self.module_pgs.values()

    def items(self):
        return self.module_pgs.items()

    def __repr__(self):
        inner = ", ".join(f"{k}: {v!r}" for k, v in self.module_pgs.items())
        return f"MultiModuleProcessGroupCollection({{{inner}}}, language_model={self.language_model_module_name!r})"
