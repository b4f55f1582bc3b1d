"""``fully_shard(module, optimizer)``: the two-line API (reference ``megatron_fsdp/fully_shard.py:1-803``).

    model, optimizer = fully_shard(model, optimizer, fsdp_unit_modules=[TransformerLayer], zero_dp_strategy=3)
    loss = model(x).sum(); loss.backward(); optimizer.step(); optimizer.zero_grad()

The returned optimizer is the one passed in, with its parameter groups re-pointed at the fp32 shard parameters and its
``step`` / ``zero_grad`` wrapped so that gradient finalisation before and weight installation after the step happen by
themselves (``sync_model_each_microbatch`` semantics)."""
from __future__ import annotations

import types
from typing import Optional, Sequence, Tuple, Union

import torch

from .megatron_fsdp import MegatronFSDP

_STRATEGY = {0: "no_shard", 1: "optim", 2: "optim_grads", 3: "optim_grads_params"}


def fully_shard_model(module: torch.nn.Module, fsdp_unit_modules: Optional[Sequence[Union[type, str]]] = None, zero_dp_strategy: Union[int, str] = 3,
                      dp_shard_group=None, dp_outer_group=None, expert_dp_group=None, ddp_config=None, preserve_fp32_weights: bool = True,
                      grad_reduce_in_fp32: bool = True, overlap_grad_reduce: bool = True, overlap_param_gather: bool = True,
                      fsdp_double_buffer: bool = False, device=None, **kw) -> MegatronFSDP:
    strategy = _STRATEGY.get(zero_dp_strategy, zero_dp_strategy)
    types_ = []
    for t in fsdp_unit_modules or ():
        if isinstance(t, str):                               # "pkg.mod.Class" paths, as accepted by the reference
            mod, _, cls = t.rpartition(".")
            t = getattr(__import__(mod, fromlist=[cls]), cls)
        types_.append(t)
    return MegatronFSDP(module, ddp_config, types_, strategy, dp_shard_group, expert_dp_group, dp_outer_group, preserve_fp32_weights, grad_reduce_in_fp32,
                        overlap_param_gather, overlap_grad_reduce, allocator=("fixed" if fsdp_double_buffer else "auto"), device=device, **kw)


def fully_shard_optimizer(model: MegatronFSDP, optimizer: torch.optim.Optimizer) -> torch.optim.Optimizer:
    """Swap the optimizer's parameters for the shard parameters (hyper-parameters of the first group are kept) and hook
    ``step`` / ``zero_grad``."""
    shards = model.optimizer_parameters()
    base = {k: v for k, v in optimizer.param_groups[0].items() if k != "params"}
    optimizer.param_groups.clear()
    optimizer.state.clear()
    optimizer.add_param_group({"params": shards, **base})
    inner_step, inner_zero = optimizer.step, optimizer.zero_grad

    def step(self, *a, **k):
        model.finish_grad_sync()
        out = inner_step(*a, **k)
        model.install_optimized_model_weights()
        return out

    def zero_grad(self, *a, **k):
        model.zero_grad_buffer()
        return inner_zero(*a, **k)

    optimizer.step = types.MethodType(step, optimizer)
    optimizer.zero_grad = types.MethodType(zero_grad, optimizer)
    return optimizer


def fully_shard(module: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, **kw) -> Tuple[MegatronFSDP, Optional[torch.optim.Optimizer]]:
    model = fully_shard_model(module, **kw)
    return model, (fully_shard_optimizer(model, optimizer) if optimizer is not None else None)
