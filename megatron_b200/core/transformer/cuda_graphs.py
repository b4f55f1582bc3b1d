"""CUDA-graph capture of transformer layers (reference ``transformer/cuda_graphs.py`` — ``CudaGraphManager`` :1854,
record/replay of autograd nodes :786-945).

B200-first rationale: a Llama-class layer launches ~40 kernels of 5-200 µs; at TP=8 the per-layer GPU time is
< 1 ms and the Python/launch overhead becomes visible.  One graph pair per (layer, input signature) replays forward
and backward with two ``cudaGraphLaunch`` calls.

Design (different from the reference's global record-then-capture pass):

* ``CudaGraphManager`` is attached to a layer (``GraphableMegatronModule``).  The first ``warmup`` calls run eagerly,
  then forward and backward are captured as TWO graphs sharing one private memory pool (``torch.cuda.make_graphed_callables``
  does the capture: static input/output/grad buffers, warm-up on a side stream, autograd wiring); the manager adds what the
  layer API needs around it — keyword tensor arguments, non-tensor outputs (``(hidden, None)``), one capture per input
  signature (shape / dtype / requires_grad / training flag), and a permanent eager fallback.
* The graphed callable is an autograd Function, so a graphed layer composes with eager neighbours, pipeline schedules
  and activation recompute of *other* layers.
* RNG: dropout inside a captured region uses PyTorch's graph-safe philox offsets.
* On CPU (unit tests) and when capture fails (data-dependent control flow, e.g. MoE token drop), the manager
  permanently falls back to eager execution for that layer and records the reason in ``self.fallback_reason``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

_POOL = None


def _shared_pool():
    global _POOL
    if _POOL is None:
        _POOL = torch.cuda.graph_pool_handle()
    return _POOL


def is_graph_capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _flatten(out):
    """→ (list of tensors, rebuild(list) -> original structure)."""
    if isinstance(out, torch.Tensor):
        return [out], lambda xs: xs[0]
    if isinstance(out, (tuple, list)):
        idx = [i for i, o in enumerate(out) if isinstance(o, torch.Tensor)]
        const = list(out)

        def rebuild(xs):
            r = list(const)
            for i, x in zip(idx, xs):
                r[i] = x
            return tuple(r)

        return [out[i] for i in idx], rebuild
    raise TypeError(f"unsupported layer output type {type(out)}")


class _Shim(torch.nn.Module):
    """Positional-tensors-in / tensors-out view of a layer call, owning the layer's parameters (what the capture API expects)."""

    def __init__(self, inner: torch.nn.Module, fn):
        super().__init__()
        self.inner = inner
        self._fn = fn

    def forward(self, *xs):
        tensors, _ = _flatten(self._fn(*xs))
        return tuple(tensors)


class _Captured:
    """One (fwd graph, bwd graph) pair for a fixed input signature."""

    def __init__(self, module: torch.nn.Module, fn, sample_args):
        with torch.no_grad():
            _, self.rebuild = _flatten(fn(*sample_args))  # learn the output structure (which slots are tensors)
        samples = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in sample_args)
        self.graphed = torch.cuda.make_graphed_callables(_Shim(module, fn), samples, num_warmup_iters=3, pool=_shared_pool())

    def __call__(self, *flat_in):
        outs = self.graphed(*flat_in)
        return self.rebuild(list(outs) if isinstance(outs, (tuple, list)) else [outs])


class CudaGraphManager:
    """Callable attached to a layer: ``manager(layer, args, kwargs)`` replays (or captures, or runs eagerly)."""

    def __init__(self, config=None, warmup_steps: Optional[int] = None, share_pool: bool = True, vp_stage=None):
        self.config = config
        self.vp_stage = vp_stage
        self.warmup = warmup_steps if warmup_steps is not None else getattr(config, "cuda_graph_warmup_steps", 3)
        self.calls = 0
        self.captured: Dict[Any, _Captured] = {}
        self.fallback_reason: Optional[str] = None
        self.share_pool = share_pool

    def should_graph(self, module, args, kwargs) -> bool:
        """Graph only calls whose tensor arguments are all on the GPU and that happen outside a capture."""
        if self.fallback_reason is not None or not torch.cuda.is_available() or is_graph_capturing():
            return False
        if kwargs.get("inference_context") is not None or kwargs.get("inference_params") is not None:
            return False  # KV-cache growth changes shapes every step
        return all(isinstance(a, torch.Tensor) and a.is_cuda for a in args)

    @staticmethod
    def _signature(args, training: bool):
        return (training,) + tuple((tuple(a.shape), a.dtype, a.requires_grad) if isinstance(a, torch.Tensor) else a for a in args)

    def __call__(self, module: torch.nn.Module, args: tuple, kwargs: dict):
        tensor_kw = {k: v for k, v in kwargs.items() if isinstance(v, torch.Tensor)}
        if not self.should_graph(module, args, kwargs):
            return module._eager_forward(*args, **kwargs)
        self.calls += 1
        if self.calls <= self.warmup:
            return module._eager_forward(*args, **kwargs)
        names = list(tensor_kw)
        flat_in = tuple(args) + tuple(tensor_kw[n] for n in names)
        const_kw = {k: v for k, v in kwargs.items() if k not in tensor_kw}
        sig = self._signature(flat_in, module.training) + tuple(names)
        cap = self.captured.get(sig)
        if cap is None:
            n_pos = len(args)

            def fn(*xs):
                return module._eager_forward(*xs[:n_pos], **dict(zip(names, xs[n_pos:])), **const_kw)

            try:
                torch.cuda.synchronize()
                cap = _Captured(module, fn, flat_in)
            except Exception as e:  # capture is best-effort: keep training eagerly
                self.fallback_reason = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
                return module._eager_forward(*args, **kwargs)
            self.captured[sig] = cap
        return cap(*flat_in)


class GraphableMixin:
    """Mixin for layers: ``forward`` routes through the manager when ``config.enable_cuda_graph`` is set."""

    def _init_cuda_graph(self, config):
        self.cudagraph_manager = CudaGraphManager(config) if getattr(config, "enable_cuda_graph", False) else None

    def __call__(self, *args, **kwargs):
        mgr = getattr(self, "cudagraph_manager", None)
        if mgr is None:
            return super().__call__(*args, **kwargs)
        return mgr(self, args, kwargs)

    def _eager_forward(self, *args, **kwargs):
        return torch.nn.Module.__call__(self, *args, **kwargs)


def graph_module(module: torch.nn.Module, warmup_steps: int = 3) -> torch.nn.Module:
    """Wrap any module instance so its calls go through a ``CudaGraphManager`` (functional form of the mixin)."""
    mgr = CudaGraphManager(warmup_steps=warmup_steps)
    cls = type(module)

    class _Graphed(cls):  # type: ignore[misc, valid-type]
        def __call__(self, *a, **k):
            return mgr(self, a, k)

        def _eager_forward(self, *a, **k):
            return cls.__call__(self, *a, **k)

    module.__class__ = _Graphed
    module.cudagraph_manager = mgr
    return module


def __getattr__(name):
    # the whole-step graph lives in ``core/full_cuda_graph.py``; kept importable from here (lazy: that module imports this one)
    if name == "FullCudaGraphWrapper":
        from ..full_cuda_graph import FullCudaGraphWrapper
        return FullCudaGraphWrapper
    raise AttributeError(name)
