"""Generalised tensor parallelism with weight rematerialisation (GTP) — reference ``tensor_parallel/generalized_tensor_parallelism.py`` (+ ``gtp_api.py``,
``gtp_cuda_graphs.py``, ``gtp_symmetric_memory.py``): per-WEIGHT ZeRO-3.  Each GEMM weight is additionally sharded ``1/R`` along its out-features over
the *remat group* (R adjacent data-parallel ranks, ``parallel_state.get_gtp_weight_remat_group``); the full weight exists only around the GEMMs that
use it — all-gathered before the forward GEMM, dropped, all-gathered again before the backward GEMMs — and the weight gradient is reduce-scattered
back to the shard.  Activations and the math are untouched, so GTP composes with TP / SP / PP and with the fused TP kernels.

How it is built here (different from the reference's TE-integrated buffers):

* ``convert_linear_to_gtp(linear)`` replaces the module's ``weight`` Parameter by ``weight_shard`` and installs two hooks.  The forward pre-hook
  materialises the full weight through ``_GTPGather`` (autograd: all-gather forward / reduce-scatter backward) and exposes it as ``module.weight``,
  so ANY linear implementation runs unchanged.  The forward hook frees the gathered storage (``untyped_storage().resize_(0)``) and registers a
  gradient hook on the layer output that RE-GATHERS into the same storage right before the layer's backward runs — the rematerialisation.
* ``GTPPrefetcher`` chains the converted modules in execution order: materialising module *i* also launches the asynchronous all-gather of module
  *i+1* (the re-gather of *i-1* in backward), so the gather is hidden behind the neighbouring GEMM and at most two full weights are resident,
  independent of depth.  Re-gathers write through an alias tensor with its own version counter, so autograd's saved-tensor check stays quiet.
* The shard Parameter is tagged ``gtp_sharded``; ``DistributedDataParallel`` reduces its gradient over the orthogonal replica group only.

On one NVSwitch box the gathers are NVLink multicast traffic; with ``R ≤ 8`` a 4096 x 14336 bf16 weight (117 MB) gathers in ≈ 0.15 ms at 800 GB/s,
well under the ≈ 1 ms GEMM it feeds — which is why prefetching one module ahead is sufficient.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import parallel_state as ps


def _gather_into(full: torch.Tensor, shard: torch.Tensor, group, async_op: bool = False):
    return dist.all_gather_into_tensor(full, shard.contiguous(), group=group, async_op=async_op)


class _GTPGather(torch.autograd.Function):
    """full[R * n, k] = all_gather(shard[n, k]);  backward: reduce-scatter (sum) of the full-weight gradient."""

    @staticmethod
    def forward(ctx, shard, group, out):
        ctx.group = group
        if out is None:
            out = shard.new_empty((shard.shape[0] * dist.get_world_size(group),) + tuple(shard.shape[1:]))
            _gather_into(out, shard.detach(), group)
        return out.view_as(out)

    @staticmethod
    def backward(ctx, g):
        r = dist.get_world_size(ctx.group)
        gs = g.new_empty((g.shape[0] // r,) + tuple(g.shape[1:]))
        dist.reduce_scatter_tensor(gs, g.contiguous(), group=ctx.group)
        return gs, None, None


def _alias(store, like: torch.Tensor) -> torch.Tensor:
    """A second tensor over ``store`` with its own autograd version counter: collectives may refill a weight that autograd saved for backward
    without tripping the "modified by an in-place operation" check (same trick as activation recompute into a released output)."""
    return torch.empty(0, dtype=like.dtype, device=like.device).set_(store, 0, like.shape, like.stride())


class _Outstanding:
    """A gathered weight whose storage has been released after forward and must be whole again for backward."""

    __slots__ = ("full", "store", "nbytes", "work")

    def __init__(self, full):
        self.full, self.store, self.nbytes, self.work = full, full.untyped_storage(), full.untyped_storage().size(), None


class GTPPrefetcher:
    """Execution-ordered chain of GTP modules.  Materialising module ``i`` launches the asynchronous gather of module ``i+1`` (forward) or the
    re-gather of module ``i-1`` (backward), so at most two full weights are resident at any time and the gather hides behind the neighbouring GEMM."""

    def __init__(self, group=None, prefetch: bool = True):
        self.group = group if group is not None else ps.get_gtp_weight_remat_group()
        self.modules: List[torch.nn.Module] = []
        self.prefetch = prefetch
        self._fwd: Dict[int, Tuple[torch.Tensor, object]] = {}        # module index -> (full tensor being gathered, work)
        self.stats = dict(gathers=0, prefetch_hits=0, regathers=0, regather_prefetch_hits=0)

    def register(self, module) -> int:
        self.modules.append(module)
        module._gtp_outstanding = []
        return len(self.modules) - 1

    # ---- forward ----
    def _launch_fwd(self, idx: int) -> None:
        if not (0 <= idx < len(self.modules)) or idx in self._fwd:
            return
        shard = self.modules[idx].weight_shard
        full = shard.new_empty((shard.shape[0] * dist.get_world_size(self.group),) + tuple(shard.shape[1:]))
        self._fwd[idx] = (full, _gather_into(full, shard.detach(), self.group, async_op=True))
        self.stats["gathers"] += 1

    def forward_weight(self, idx: int) -> torch.Tensor:
        if idx in self._fwd:
            self.stats["prefetch_hits"] += 1
        else:
            self._launch_fwd(idx)
        full, work = self._fwd.pop(idx)
        if work is not None:
            work.wait()
        if self.prefetch:
            self._launch_fwd(idx + 1)
        return full

    # ---- backward ----
    def _launch_bwd(self, rec: _Outstanding, shard: torch.Tensor) -> None:
        if rec.work is not None or rec.store.size() != 0:
            return
        rec.store.resize_(rec.nbytes)
        rec.work = _gather_into(_alias(rec.store, rec.full), shard.detach(), self.group, async_op=True)
        self.stats["regathers"] += 1

    def backward_weight(self, idx: int, rec: _Outstanding) -> None:
        mod = self.modules[idx]
        if rec.work is not None:
            self.stats["regather_prefetch_hits"] += 1
        self._launch_bwd(rec, mod.weight_shard)
        if rec.work is not None:
            rec.work.wait()
            rec.work = None
        if self.prefetch and idx - 1 >= 0:
            prev = self.modules[idx - 1]
            if prev._gtp_outstanding:
                self._launch_bwd(prev._gtp_outstanding[-1], prev.weight_shard)

    def reset(self) -> None:
        for _, w in self._fwd.values():
            if w is not None:
                w.wait()
        self._fwd.clear()


def convert_linear_to_gtp(module: torch.nn.Module, prefetcher: Optional[GTPPrefetcher] = None, group=None, attr: str = "weight", expert: bool = False) -> torch.nn.Module:
    """Shard ``module.<attr>`` along dim 0 (``[out, in]`` with out % R == 0, or the expert axis of a grouped ``[L, …]`` weight) over the remat group,
    in place.  Works for ``ColumnParallelLinear`` / ``RowParallelLinear`` / ``torch.nn.Linear`` / ``GroupedMLP`` — anything that reads
    ``self.<attr>`` in ``forward``.  ``expert=True`` tags the shard so DDP reduces it over the expert replicas."""
    if prefetcher is None:
        prefetcher = GTPPrefetcher(group=group, prefetch=False)
    group = prefetcher.group
    r, rank = dist.get_world_size(group), dist.get_rank(group)
    w = getattr(module, attr)
    assert w.shape[0] % r == 0, f"out-features {w.shape[0]} not divisible by the GTP remat size {r}"
    n = w.shape[0] // r
    shard = torch.nn.Parameter(w.detach()[rank * n : (rank + 1) * n].clone(), requires_grad=w.requires_grad)
    for tag in ("tensor_model_parallel", "partition_dim", "partition_stride", "sequence_parallel", "allreduce", "is_embedding_or_output_parameter"):
        if hasattr(w, tag):
            setattr(shard, tag, getattr(w, tag))
    shard.gtp_sharded = True
    shard.gtp_expert = expert
    shard.gtp_full_shape = tuple(w.shape)
    del module._parameters[attr]
    shard_name = f"{attr}_shard"
    module.register_parameter(shard_name, shard)
    module.gtp_group = group
    module.gtp_prefetcher = prefetcher
    # the prefetcher reads ``.weight_shard`` of whatever it registered: give every converted attribute its own handle
    handle = module if attr == "weight" else _AttrHandle(module, shard_name)
    index = prefetcher.register(handle)
    if attr == "weight":
        module.gtp_index = index
    module.__dict__.setdefault("gtp_attrs", {})[attr] = index

    def pre_hook(mod, args):
        buf = mod.gtp_prefetcher.forward_weight(index)
        object.__setattr__(mod, attr, _GTPGather.apply(getattr(mod, shard_name), mod.gtp_group, buf))

    def post_hook(mod, args, output):
        full = mod.__dict__.pop(attr)
        out = output[0] if isinstance(output, tuple) else output
        if not (torch.is_grad_enabled() and isinstance(out, torch.Tensor) and out.requires_grad):
            return
        rec = _Outstanding(full)
        rec.store.resize_(0)                 # the full weight is gone until this layer's backward needs it
        handle._gtp_outstanding.append(rec)

        def regather(grad):                  # runs right before this layer's backward node
            mod.gtp_prefetcher.backward_weight(index, rec)
            if rec in handle._gtp_outstanding:
                handle._gtp_outstanding.remove(rec)
            return grad

        out.register_hook(regather)

    module.register_forward_pre_hook(pre_hook)
    module.register_forward_hook(post_hook)
    return module


class _AttrHandle:
    """What the prefetcher needs from a registered weight (``weight_shard``, its own ``_gtp_outstanding`` list) for an attribute that is not called
    ``weight`` — a module with two converted attributes (``GroupedMLP.weight1 / weight2``) gets two independent links in the chain."""

    def __init__(self, module, shard_name):
        self._m, self._n = module, shard_name
        self._gtp_outstanding = []

    @property
    def weight_shard(self):
        return getattr(self._m, self._n)


def apply_expert_gtp(model: torch.nn.Module, prefetch: bool = True) -> "GTPPrefetcher":
    """Expert-side GTP (reference EGTP groups, ``parallel_state`` ``expert_gtp_remat_size``): shard every expert weight over the expert remat group.
    ``GroupedMLP`` shards its stacked ``weight1 / weight2`` along the local-expert axis (``L % R == 0``), ``SequentialMLP`` experts shard each linear's
    out-features like the dense case."""
    from ..transformer.moe.experts import GroupedMLP, SequentialMLP

    group = ps.get_expert_gtp_weight_remat_group()
    assert group is not None, "initialize_model_parallel(expert_gtp_remat_size=R) first"
    pre = GTPPrefetcher(group=group, prefetch=prefetch)
    r = dist.get_world_size(group)
    for _, mod in list(model.named_modules()):
        if isinstance(mod, GroupedMLP):
            assert mod.weight1.shape[0] % r == 0, f"{mod.weight1.shape[0]} local experts cannot be sharded over an expert remat group of {r}"
            convert_linear_to_gtp(mod, pre, attr="weight1", expert=True)
            convert_linear_to_gtp(mod, pre, attr="weight2", expert=True)
        elif isinstance(mod, SequentialMLP):
            for e in mod.local_experts:
                for lin in (e.linear_fc1, e.linear_fc2):
                    if lin.weight.shape[0] % r == 0:
                        convert_linear_to_gtp(lin, pre, expert=True)
    return pre


def apply_gtp(model: torch.nn.Module, min_numel: int = 1 << 20, prefetch: bool = True, predicate=None) -> GTPPrefetcher:
    """Convert every large 2-D linear weight of ``model`` (in registration = execution order) and return the prefetcher that chains them
    (reference ``gtp_api.py``).  ``predicate(name, module)`` overrides the default choice (TP linears and ``nn.Linear`` with ≥ ``min_numel`` weights)."""
    from .layers import ColumnParallelLinear, RowParallelLinear

    pre = GTPPrefetcher(prefetch=prefetch)
    r = dist.get_world_size(pre.group)
    for name, mod in list(model.named_modules()):
        w = getattr(mod, "weight", None)
        if not isinstance(w, torch.nn.Parameter) or w.dim() != 2 or w.shape[0] % r:
            continue
        ok = predicate(name, mod) if predicate is not None else (isinstance(mod, (ColumnParallelLinear, RowParallelLinear, torch.nn.Linear)) and w.numel() >= min_numel)
        if ok:
            convert_linear_to_gtp(mod, pre)
    return pre


def gather_gtp_state_dict(model: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """Full (un-sharded) weights of all GTP modules, keyed like the original ``weight`` entries — for export / checkpoint conversion."""
    out = {}
    for name, mod in model.named_modules():
        if hasattr(mod, "weight_shard") and hasattr(mod, "gtp_group"):
            r = dist.get_world_size(mod.gtp_group)
            full = mod.weight_shard.new_empty((mod.weight_shard.shape[0] * r,) + tuple(mod.weight_shard.shape[1:]))
            _gather_into(full, mod.weight_shard.detach(), mod.gtp_group)
            out[f"{name}.weight"] = full
    return out
