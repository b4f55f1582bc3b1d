"""Whole-iteration CUDA graph: every micro-batch's forward AND backward of a training step captured once and replayed
(reference ``core/full_cuda_graph.py:1-267``).

Why on B200: an 8B model at TP=8 launches ~3 k kernels per step for ~300 ms of device time — the host barely keeps up, and any
hiccup (Python GC, a logging call) stalls the device.  One ``cudaGraphLaunch`` per step removes the host from the loop.

Requirements on the captured function (the schedules satisfy them when ``--full-cuda-graph`` is set): static shapes, no host
synchronisation, no data-dependent Python control flow, gradients accumulated into persistent ``main_grad`` buffers, RNG through
the graph-safe tracker.  The optimizer step stays OUTSIDE the graph (its scalars — lr, loss scale, clip coefficient — change
per step).  Input batches are copied into static device buffers by ``StaticBufferLoader`` before every replay."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

_SHARED_POOL = None
_SHARED_STREAM = None


def get_shared_graph_pool():
    """One memory pool for every graph of the process: graphs that never run concurrently (train / eval, per-layer graphs) reuse
    each other's activations memory."""
    global _SHARED_POOL
    if _SHARED_POOL is None:
        _SHARED_POOL = torch.cuda.graph_pool_handle()
    return _SHARED_POOL


get_graph_pool = get_shared_graph_pool


def get_shared_capture_stream():
    """Captures must not run on the legacy default stream; all captures share one side stream."""
    global _SHARED_STREAM
    if _SHARED_STREAM is None:
        _SHARED_STREAM = torch.cuda.Stream()
    return _SHARED_STREAM


def copy_tensors_in_struct(dst: Any, src: Any) -> Any:
    """``dst <- src`` for every tensor of two identically shaped nests (dict / list / tuple); non-tensor leaves of ``dst`` are
    replaced by ``src``'s.  Shapes must match: a static buffer cannot follow a batch that changed shape."""
    if torch.is_tensor(dst):
        if not torch.is_tensor(src) or dst.shape != src.shape:
            raise ValueError(f"static buffer {tuple(dst.shape)} cannot take {tuple(src.shape) if torch.is_tensor(src) else type(src).__name__}: "
                             "shapes must stay fixed under a full-iteration CUDA graph")
        dst.copy_(src, non_blocking=True)
        return dst
    if isinstance(dst, dict):
        if not isinstance(src, dict) or dst.keys() != src.keys():
            raise ValueError("batch structure changed under a full-iteration CUDA graph")
        for k in dst:
            dst[k] = copy_tensors_in_struct(dst[k], src[k])
        return dst
    if isinstance(dst, (list, tuple)):
        if not isinstance(src, (list, tuple)) or len(dst) != len(src):
            raise ValueError("batch structure changed under a full-iteration CUDA graph")
        out = [copy_tensors_in_struct(d, s) for d, s in zip(dst, src)]
        if isinstance(dst, list):
            dst[:] = out
            return dst
        return type(dst)(out)
    return src


def clone_tensors_in_struct(src: Any, device=None) -> Any:
    """Deep copy of a nest with every tensor cloned onto ``device`` (the static buffers are born here)."""
    if torch.is_tensor(src):
        return src.to(device, copy=True) if device is not None else src.clone()
    if isinstance(src, dict):
        return {k: clone_tensors_in_struct(v, device) for k, v in src.items()}
    if isinstance(src, (list, tuple)):
        return type(src)(clone_tensors_in_struct(v, device) for v in src)
    return src


class StaticBufferLoader:
    """Fixed-address copies of the micro-batches of a step, per stage (``training`` / ``validation``): the graph reads these.

    ``load(stage, microbatch, batch)`` creates the buffer on first use and copies into it afterwards (asynchronously on the
    current stream; host batches should be pinned).  The iterator handed to the captured function yields the static buffers."""

    def __init__(self, device=None):
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self.static_buffers: Dict[str, List[Any]] = {"training": [], "validation": []}

    def load(self, stage: str, microbatch: int, batch: Any) -> Any:
        bufs = self.static_buffers.setdefault(stage, [])
        if microbatch == len(bufs):
            bufs.append(clone_tensors_in_struct(batch, self.device))
        elif microbatch < len(bufs):
            bufs[microbatch] = copy_tensors_in_struct(bufs[microbatch], batch)
        else:
            raise IndexError(f"micro-batch {microbatch} loaded before {len(bufs)}")
        return bufs[microbatch]

    def load_step(self, stage: str, data_iterator, num_microbatches: int) -> List[Any]:
        its = data_iterator if isinstance(data_iterator, (list, tuple)) else [data_iterator]
        # virtual pipeline: one iterator per model chunk, all yielding the same micro-batches -> read the first, share the buffers
        return [self.load(stage, i, next(its[0])) for i in range(num_microbatches)]

    class _Iter:
        def __init__(self, bufs):
            self.bufs, self.i = bufs, 0

        def __iter__(self):
            return self

        def __next__(self):
            b = self.bufs[self.i % len(self.bufs)]
            self.i += 1
            return b

    def iterator(self, stage: str, like=None):
        it = self._Iter(self.static_buffers[stage])
        if isinstance(like, (list, tuple)):
            return [self._Iter(self.static_buffers[stage]) for _ in like]
        return it


class FullCudaGraphWrapper:
    """``wrapped(data_iterator=…, num_microbatches=…, forward_only=…, **kw)`` with the signature of a forward-backward
    function: eager for ``cuda_graph_warmup_steps`` calls per stage (allocator warm-up, lazy initialisations), then captured on
    the shared stream into the shared pool, then replayed.  Returns the STATIC result tensors of the capture (losses); read
    them before the next replay.  Without CUDA it runs the function eagerly on the static buffers (so the data path is
    testable on CPU)."""

    def __init__(self, forward_backward_func, cuda_graph_warmup_steps: int = 1, device=None):
        self.fn = forward_backward_func
        self.warmup = cuda_graph_warmup_steps
        self.loader = StaticBufferLoader(device)
        self.calls = {"training": 0, "validation": 0}
        self.graph: Dict[str, "torch.cuda.CUDAGraph"] = {}
        self.result: Dict[str, Any] = {}

    def reset(self) -> None:
        """Drop the graphs (model structure, shapes or parallel layout changed)."""
        self.graph.clear()
        self.result.clear()
        self.calls = {k: 0 for k in self.calls}
        self.loader.static_buffers = {"training": [], "validation": []}

    def __call__(self, *, data_iterator, num_microbatches: int, forward_only: bool = False, **kwargs):
        stage = "validation" if forward_only else "training"
        self.loader.load_step(stage, data_iterator, num_microbatches)
        it = self.loader.iterator(stage, like=data_iterator)
        self.calls[stage] += 1
        if not torch.cuda.is_available() or self.calls[stage] <= self.warmup:
            return self.fn(data_iterator=it, num_microbatches=num_microbatches, forward_only=forward_only, **kwargs)
        if stage not in self.graph:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            stream = get_shared_capture_stream()
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, pool=get_shared_graph_pool(), stream=stream):
                self.result[stage] = self.fn(data_iterator=it, num_microbatches=num_microbatches, forward_only=forward_only, **kwargs)
            torch.cuda.current_stream().wait_stream(stream)
            self.graph[stage] = g
        self.graph[stage].replay()
        return self.result[stage]
