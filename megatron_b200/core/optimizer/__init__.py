"""Optimizer factory (reference ``optimizer/__init__.py:993`` ``get_megatron_optimizer``)."""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional, Tuple

import torch

from .. import parallel_state as ps
from ..utils import get_model_config, get_pg_size
from .distrib_optimizer import DistributedOptimizer
from .grad_scaler import ConstantGradScaler, DynamicGradScaler
from .optimizer import ChainedOptimizer, Float16OptimizerWithFloat16Params, FP32Optimizer, MegatronOptimizer
from .optimizer_config import OptimizerConfig

logger = logging.getLogger(__name__)


def _match_override(name: str, p, overrides):
    """``overrides``: ``{pattern | callable(name, param) | ParamKey-like with .name / .attr: {"lr_mult"|"max_lr"|"min_lr"|"wd_mult"|"weight_decay"|...: value}}`` — first match wins
    (reference ``get_megatron_optimizer(config_overrides=...)``: per-parameter optimizer settings keyed by name globs / attributes)."""
    import fnmatch

    for key, ov in (overrides or {}).items():
        if callable(key):
            hit = key(name, p)
        elif isinstance(key, str):
            hit = fnmatch.fnmatch(name, key)
        else:
            names = getattr(key, "name", None) or ()
            attrs = getattr(key, "attr", None) or ()
            names = (names,) if isinstance(names, str) else names
            attrs = (attrs,) if isinstance(attrs, str) else attrs
            hit = any(fnmatch.fnmatch(name, n) for n in names) or any(getattr(p, a, False) for a in attrs)
        if hit:
            return tuple(sorted(ov.items()))
    return ()


def _get_param_groups(model_chunks: List, no_weight_decay_cond: Optional[Callable], scale_lr_cond: Optional[Callable], lr_mult: float,
                      lr: float, min_lr: float, decoupled_lr: Optional[float], decoupled_min_lr: Optional[float], default_wd: float = 0.01,
                      config_overrides=None, mup_width_mult: Optional[float] = None) -> List[Dict]:
    """Bucket params by (wd_mult, lr_mult, is_expert_parallel, is_decoupled_lr, overrides).
    Default rule: biases and 1-D tensors (norm gains) get no weight decay.  ``mup_width_mult`` (= hidden / base hidden) applies the µP Adam rule: matrix-like
    hidden weights train with ``lr / width_mult``, vector-like parameters (embeddings, norms, biases) keep the base learning rate."""
    use_decoupled = decoupled_lr is not None
    buckets: Dict[Tuple, List] = {}
    for chunk in model_chunks:
        for name, p in chunk.named_parameters():
            if not p.requires_grad:
                continue
            is_expert = not getattr(p, "allreduce", True)
            if no_weight_decay_cond is not None:
                no_wd = no_weight_decay_cond(name, p)
            else:
                no_wd = name.endswith(".bias") or p.dim() == 1
            scale_lr = scale_lr_cond(name, p) if scale_lr_cond is not None else False
            wd_mult = 0.0 if no_wd else 1.0
            lm = lr_mult if scale_lr else 1.0
            dec = use_decoupled and getattr(p, "is_embedding_or_output_parameter", False)
            mup_hidden = bool(mup_width_mult and mup_width_mult != 1.0 and p.dim() >= 2 and not getattr(p, "is_embedding_or_output_parameter", False) and "embedding" not in name)
            if mup_hidden:
                lm = lm / mup_width_mult
            ov = _match_override(name, p, config_overrides)
            buckets.setdefault((wd_mult, lm, is_expert, dec, ov, mup_hidden), []).append(p)
    groups = []
    for (wd_mult, lm, is_expert, dec, ov, mup_hidden), params in buckets.items():
        ovd = dict(ov)
        wd_mult, lm = ovd.pop("wd_mult", wd_mult), ovd.pop("lr_mult", lm)
        max_lr = ovd.pop("max_lr", decoupled_lr if dec else lr)
        g = dict(params=params, wd_mult=wd_mult, lr_mult=lm, is_expert_parallel=is_expert, is_decoupled_lr=dec,
                 max_lr=max_lr, min_lr=ovd.pop("min_lr", decoupled_min_lr if dec and decoupled_min_lr is not None else min_lr),
                 lr=max_lr * lm if max_lr is not None else None, weight_decay=ovd.pop("weight_decay", default_wd * wd_mult))
        g.update(ovd)            # anything else (betas, eps, ...) goes straight into the group
        if mup_hidden:
            g["mup_hidden"] = True
        groups.append(g)
    return groups


def _make_scaler(config: OptimizerConfig):
    if config.loss_scale:
        return ConstantGradScaler(config.loss_scale)
    if config.fp16:
        return DynamicGradScaler(config.initial_loss_scale, config.min_loss_scale, 2.0, 0.5, config.loss_scale_window, config.hysteresis)
    return None


def _get_megatron_optimizer_based_on_param_groups(config: OptimizerConfig, model_chunks, param_groups, per_model_buffers=None,
                                                  model_parallel_group=None, data_parallel_group=None, data_parallel_group_gloo=None,
                                                  data_parallel_group_idx=0, distributed_optimizer_instance_id=0) -> MegatronOptimizer:
    for g in param_groups:
        g.setdefault("betas", (config.adam_beta1, config.adam_beta2))
        g.setdefault("eps", config.adam_eps)
    lowp = config.fp16 or config.bf16
    scaler = _make_scaler(config)
    from .emerging_optimizers import needs_whole_matrices

    dp_ws = torch.distributed.get_world_size(data_parallel_group) if (data_parallel_group is not None and torch.distributed.is_initialized()) else 1
    if needs_whole_matrices(config.optimizer) and (config.use_distributed_optimizer or getattr(config, "use_layer_wise_distributed_optimizer", False)) and dp_ws > 1:
        # Muon-class rules cannot run on ZeRO-1 flat ranges: distribute by parameter ownership instead
        from .layer_wise_optimizer import LayerWiseDistributedOptimizer

        return LayerWiseDistributedOptimizer(config, model_chunks, param_groups, data_parallel_group, scaler, model_parallel_group)
    if config.use_distributed_optimizer and not needs_whole_matrices(config.optimizer):
        opt = DistributedOptimizer(param_groups, config, scaler, None, model_chunks, per_model_buffers or {}, data_parallel_group,
                                   data_parallel_group_gloo, data_parallel_group_idx, distributed_optimizer_instance_id)
    elif lowp:
        opt = Float16OptimizerWithFloat16Params(param_groups, config, scaler)
        opt.model_chunks = model_chunks
    else:
        opt = FP32Optimizer(param_groups, config)
        opt.model_chunks = model_chunks
    if model_parallel_group is not None and not config.use_distributed_optimizer:
        opt.grad_stats_parallel_group = model_parallel_group
    return opt


def get_megatron_optimizer(config: OptimizerConfig, model_chunks: List, no_weight_decay_cond: Optional[Callable] = None,
                           scale_lr_cond: Optional[Callable] = None, lr_mult: float = 1.0, config_overrides=None,
                           use_gloo_process_groups: bool = True, pg_collection=None, dump_param_to_param_group_map=None) -> MegatronOptimizer:
    """Dense and expert-parallel parameters get separate optimizers (their data-parallel
    groups differ) chained into one."""
    mup = None
    mc = get_model_config(model_chunks[0])
    if getattr(config, "use_mup", False) and getattr(config, "mup_base_hidden_size", None):
        mup = mc.hidden_size / config.mup_base_hidden_size
    elif getattr(mc, "use_mup", False):
        mup = mc.mup_width_mult                                        # the model config carries the width multiplier (--use-mup)
    groups = _get_param_groups(model_chunks, no_weight_decay_cond, scale_lr_cond, lr_mult, config.lr, config.min_lr,
                               config.decoupled_lr, config.decoupled_min_lr, config.weight_decay, config_overrides=config_overrides, mup_width_mult=mup)
    for g in groups:
        if g.pop("mup_hidden", False) and config.optimizer == "adam" and "eps" not in g:
            g["eps"] = config.adam_eps / mup           # µP Adam: the epsilon of width-scaled matrices shrinks with their gradients (reference get_mup_config_overrides)
    dense = [g for g in groups if not g["is_expert_parallel"]]
    expert = [g for g in groups if g["is_expert_parallel"]]
    init = ps.is_initialized()
    opts = []
    dense_buffers = {i: getattr(c, "buffers", []) for i, c in enumerate(model_chunks)}
    expert_buffers = {i: getattr(c, "expert_parallel_buffers", []) for i, c in enumerate(model_chunks)}
    if dense or not expert:
        dpg = (pg_collection.dp_cp if pg_collection is not None else (ps.get_data_parallel_group(with_context_parallel=True, partial_data_parallel=True) if init else None))
        opts.append(_get_megatron_optimizer_based_on_param_groups(
            config, model_chunks, dense, dense_buffers,
            model_parallel_group=(ps.get_model_parallel_group(check_initialized=False) if init else None),
            data_parallel_group=dpg, data_parallel_group_idx=(ps.get_model_parallel_rank() if init and False else 0),
        ))
    if expert:
        opts.append(_get_megatron_optimizer_based_on_param_groups(
            config, model_chunks, expert, expert_buffers,
            model_parallel_group=(ps.get_group("tp_ep_pp", check_initialized=False) if init else None),
            data_parallel_group=(ps.get_expert_data_parallel_group() if init else None), data_parallel_group_idx=1,
        ))
    for o in opts:
        o.tp_group = ps.get_tensor_model_parallel_group(check_initialized=False) if init else None
    return ChainedOptimizer(opts)
