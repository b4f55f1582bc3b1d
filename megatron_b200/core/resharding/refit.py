"""Weight refit between a training model and a serving model (reference ``resharding/refit.py:39-495``).

    swap_model_weights(train_model, serve_model, "nvlink")          # collocated: every rank holds both
    swap_model_weights(train_model, None, "nccl", dst_rank_offset=8) # non-collocated: ranks 0-7 train, 8-15 serve

Plans are cached per (layout pair, group, offsets): RL loops call this every iteration and the plan depends only on shapes
and process-group geometry.  ``prepare_swap_model_weights`` builds the plan (and the MXFP8 transform of a quantised serving
model) ahead of time, while both models still expose plain bf16 parameters."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch
import torch.distributed as dist

from .copy_services import CopyService, GlooCopyService, NCCLCopyService, NVLinkCopyService
from .execution import execute_reshard_plan
from .planner import build_local_reshard_plan
from .transforms import MXFP8ReshardTransform, ReshardTransform
from .utils import ReshardPlan, get_refit_tensor_dict, named_persistent_buffers

RefitBackendName = str          # "nccl" | "gloo" | "nvlink" (aliases: "nvshmem", "nixl" -> "nvlink": one-sided stores over NVLink)

_ALIASES = {"nvshmem": "nvlink", "nixl": "nvlink", "nvl": "nvlink"}
_SERVICES: Dict[Tuple[str, int], CopyService] = {}
_PLANS: Dict["_PlanCacheKey", ReshardPlan] = {}


@dataclass(frozen=True)
class _PlanCacheKey:
    src_cfg: Optional[Tuple]
    dst_cfg: Optional[Tuple]
    src_sig: Optional[int]
    dst_sig: Optional[int]
    group: int
    src_rank_offset: int
    dst_rank_offset: int
    pool_index: int


def _pgs_of(core):
    """Process-group bundle of a model: its own ``pg_collection`` or the global parallel state."""
    pgc = getattr(core, "pg_collection", None)
    if pgc is not None:
        return pgc
    from .. import parallel_state as ps
    if not ps.model_parallel_is_initialized():
        return None

    class _G:
        tp = ps.get_tensor_model_parallel_group()
        pp = ps.get_pipeline_model_parallel_group()
        dp = ps.get_data_parallel_group()
        ep = ps.get_expert_model_parallel_group(check_initialized=False) if hasattr(ps, "get_expert_model_parallel_group") else None
        expt_tp = ps.get_expert_tensor_parallel_group(check_initialized=False) if hasattr(ps, "get_expert_tensor_parallel_group") else None
        expt_dp = ps.get_expert_data_parallel_group(check_initialized=False) if hasattr(ps, "get_expert_data_parallel_group") else None
    return _G


def _sizes(pgc) -> Optional[Tuple[int, int, int, int, int]]:
    if pgc is None:
        return None

    def n(g):
        if g is None:
            return 1
        return len(g) if isinstance(g, (list, tuple)) else dist.get_world_size(g)
    return tuple(n(getattr(pgc, k, None)) for k in ("tp", "pp", "ep", "dp", "expt_tp"))


def _get_config_tuple(core) -> Optional[Tuple[int, int, int, int, int]]:
    return _sizes(_pgs_of(core)) if core is not None else None


def _signature(core) -> Optional[int]:
    if core is None:
        return None
    return hash(tuple((n, tuple(t.shape), str(t.dtype)) for n, t in get_refit_tensor_dict(core).items()))


def _build_plan_cache_key(src_core, tgt_core, group, src_rank_offset, dst_rank_offset, pool_index=0) -> _PlanCacheKey:
    return _PlanCacheKey(_get_config_tuple(src_core), _get_config_tuple(tgt_core), _signature(src_core), _signature(tgt_core),
                         id(group) if group is not None else 0, src_rank_offset, dst_rank_offset, pool_index)


def get_or_create_service(backend: RefitBackendName, group=None) -> CopyService:
    name = _ALIASES.get(backend, backend)
    key = (name, id(group) if group is not None else 0)
    svc = _SERVICES.get(key)
    if svc is None:
        if name == "nccl":
            svc = NCCLCopyService(group)
        elif name == "gloo":
            svc = GlooCopyService(group)
        elif name == "nvlink":
            svc = NVLinkCopyService(group)
        else:
            raise ValueError(f"unknown refit backend '{backend}' (nccl | gloo | nvlink)")
        _SERVICES[key] = svc
    return svc


def clear_service_cache():
    for s in _SERVICES.values():
        s.close()
    _SERVICES.clear()


def clear_plan_cache():
    _PLANS.clear()


def clear_all_caches():
    clear_plan_cache()
    clear_service_cache()


def _unwrap(m):
    while m is not None and hasattr(m, "module") and isinstance(getattr(m, "module"), torch.nn.Module):
        m = m.module
    return m


def _unwrap_model_cores(src_model, target_model):
    """Strip DDP / Float16Module wrappers and model-chunk lists; read the expert count off whichever side exists."""
    def one(m):
        if isinstance(m, (list, tuple)):
            assert len(m) == 1, "refit of interleaved (virtual pipeline) chunk lists: pass the chunks one by one"
            m = m[0]
        return _unwrap(m)
    s, t = one(src_model), one(target_model)
    cfg = getattr(s if s is not None else t, "config", None)
    return s, t, getattr(cfg, "num_moe_experts", None) if cfg is not None else None


def _build_or_get_plan(src_core, tgt_core, num_experts, group, src_rank_offset, dst_rank_offset, pool_index: int = 0) -> ReshardPlan:
    key = _build_plan_cache_key(src_core, tgt_core, group, src_rank_offset, dst_rank_offset, pool_index)
    plan = _PLANS.get(key)
    if plan is None:
        plan = build_local_reshard_plan(src_core, tgt_core, _pgs_of(src_core) if src_core is not None else None,
                                        _pgs_of(tgt_core) if tgt_core is not None else None, num_experts, group, src_rank_offset, dst_rank_offset)
        _PLANS[key] = plan
    return plan


def _needs_mxfp8_conversion(model) -> bool:
    cfg = getattr(_unwrap(model), "config", None) if model is not None else None
    return bool(cfg is not None and getattr(cfg, "transformer_impl", None) == "inference_optimized" and getattr(cfg, "fp8_recipe", None) == "mxfp8")


def _setup_mxfp8_transform_on_plan(plan: ReshardPlan, target_model) -> None:
    """Quantised linears of the serving model expose ``mxfp8_refit_buffers() -> (payload, scales)`` and
    ``refresh_mxfp8_layout()``; their ``weight`` names become the convertible set."""
    if plan.transform is not None or not _needs_mxfp8_conversion(target_model):
        return
    core = _unwrap(target_model)
    buffers, refresh = {}, {}
    for mod_name, mod in core.named_modules():
        fn = getattr(mod, "mxfp8_refit_buffers", None)
        if callable(fn):
            name = f"{mod_name}.weight" if mod_name else "weight"
            buffers[name] = fn()
            refresh[name] = getattr(mod, "refresh_mxfp8_layout", lambda: None)
    if buffers:
        plan.transform = MXFP8ReshardTransform(buffers, refresh=lambda n: refresh[n]())


def prepare_swap_model_weights(src_model, target_model, group=None, src_rank_offset: int = 0, dst_rank_offset: int = 0):
    """Build and cache the plan (collective).  Call while both models still hold plain parameters (reference ``refit.py:295``)."""
    s, t, ne = _unwrap_model_cores(src_model, target_model)
    plan = _build_or_get_plan(s, t, ne, group, src_rank_offset, dst_rank_offset)
    _setup_mxfp8_transform_on_plan(plan, target_model)
    return plan


def _harmonize_buffer_dtypes(plan: ReshardPlan, src_core, tgt_core, group=None):
    """Persistent buffers (router expert bias, ...) may be fp32 on one side and bf16 on the other after ``Float16Module``;
    the wire dtype is the sender's and the receiver converts, so nothing has to be cast in place — only recorded once for
    reports (reference ``refit.py:414`` casts the destination buffers instead)."""
    if plan.buffer_dtypes is None and tgt_core is not None:
        plan.buffer_dtypes = {n: b.dtype for n, b in named_persistent_buffers(tgt_core)}


def reshard_model_weights(src_model, target_model, service: CopyService, group=None, src_rank_offset: int = 0, dst_rank_offset: int = 0,
                          transform: Optional[ReshardTransform] = None, pool_index: int = 0):
    """(src, tgt) collocated · (src, None) pure sender · (None, tgt) pure receiver · (None, None) idle participant."""
    s, t, ne = _unwrap_model_cores(src_model, target_model)
    plan = _build_or_get_plan(s, t, ne, group, src_rank_offset, dst_rank_offset, pool_index)
    _harmonize_buffer_dtypes(plan, s, t, group)
    execute_reshard_plan(plan, s, t, service=service, group=group, transform=transform)
    return plan


def swap_model_weights(src_model, target_model, refit_method: Union[RefitBackendName, CopyService] = "nccl", group=None,
                       src_rank_offset: int = 0, dst_rank_offset: int = 0, transform: Optional[ReshardTransform] = None):
    """The call an RL loop makes after every optimizer step (reference ``refit.py:342``)."""
    service = refit_method if isinstance(refit_method, CopyService) else get_or_create_service(refit_method, group)
    return reshard_model_weights(src_model, target_model, service, group, src_rank_offset, dst_rank_offset, transform)
