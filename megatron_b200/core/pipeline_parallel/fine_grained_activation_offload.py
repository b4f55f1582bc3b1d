"""Per-module activation offload to pinned host memory (reference ``pipeline_parallel/fine_grained_activation_offload.py`` —
``OffloadTensorPool`` :150, ``PipelineOffloadManager`` :443, ``ChunkOffloadHandler`` :881, interface :1496).

Mechanism here: ``torch.autograd.graph.saved_tensors_hooks``.  Inside ``with mgr.group(name)`` every tensor autograd saves for
backward (that is large enough and is not a parameter) is copied device→host on a dedicated D2H stream into a pooled pinned
buffer, and the hook returns a small ticket instead of the tensor, so the device copy dies as soon as the copy has drained.
``mgr.commit(out, name)`` closes the group; it is an identity autograd node whose *backward* — which runs right before the
group's own backward — prefetches the PREVIOUS group's tensors host→device on the H2D stream, so reload of group k-1 overlaps
the backward compute of group k.  B200 specifics: 180 GB of HBM make offload a tool for long-sequence / large-micro-batch
runs rather than a necessity, so the default threshold only offloads ≥ 1 M-element tensors and at most ``max_inflight_bytes``
are in flight on the ~55 GB/s PCIe Gen5 link.

On CPU tensors the "host copy" is a clone, which keeps the bookkeeping testable without a GPU.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Optional

import torch

_ENABLED = [True]


def fine_grained_offloading_disable_offload():
    _ENABLED[0] = False


def fine_grained_offloading_enable_offload():
    _ENABLED[0] = True


class OffloadTensorPool:
    """Pinned host buffers keyed by (dtype, numel): activations have the same shapes every micro-batch, so after the first
    step no ``cudaHostAlloc`` happens on the hot path."""

    def __init__(self, pin: bool = True):
        self.pin = pin and torch.cuda.is_available()
        self.free: Dict[tuple, List[torch.Tensor]] = {}
        self.allocated_bytes = 0

    def get(self, like: torch.Tensor) -> torch.Tensor:
        key = (like.dtype, like.numel())
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        self.allocated_bytes += like.numel() * like.element_size()
        return torch.empty(like.numel(), dtype=like.dtype, device="cpu", pin_memory=self.pin)

    def put(self, buf: torch.Tensor) -> None:
        self.free.setdefault((buf.dtype, buf.numel()), []).append(buf)


class _Ticket:
    __slots__ = ("group", "host", "shape", "stride_ok", "device", "dev", "event", "uses")

    def __init__(self, group, host, shape, device):
        self.group, self.host, self.shape, self.device = group, host, shape, device
        self.dev: Optional[torch.Tensor] = None
        self.event = None
        self.uses = 0


class _Group:
    def __init__(self, name, index):
        self.name, self.index = name, index
        self.tickets: List[_Ticket] = []
        self.prefetched = False


class _Commit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mgr, index):
        ctx.mgr, ctx.index = mgr, index
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.mgr._on_backward_reaches(ctx.index)
        return g, None, None


class FineGrainedActivationOffloadingInterface:
    def __init__(self, min_offload_numel: int = 1 << 20, pin: bool = True, max_inflight_bytes: int = 8 << 30):
        self.min_numel, self.max_inflight = min_offload_numel, max_inflight_bytes
        self.pool = OffloadTensorPool(pin)
        self.groups: List[_Group] = []
        self._cur: Optional[_Group] = None
        self._stats = dict(groups=0, tensors_offloaded=0, bytes_offloaded=0, live=0, sync_reloads=0)
        self.d2h = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.h2d = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._param_ptrs = None

    # ---- hooks ----
    def _eligible(self, t: torch.Tensor) -> bool:
        if not _ENABLED[0] or self._cur is None or not isinstance(t, torch.Tensor):
            return False
        if isinstance(t, torch.nn.Parameter) or t.numel() < self.min_numel or not t.is_contiguous() or getattr(t, "_do_not_offload", False):
            return False
        return t._base is None or not isinstance(t._base, torch.nn.Parameter)

    def _pack(self, t):
        if not self._eligible(t):
            return t
        host = self.pool.get(t)
        tk = _Ticket(self._cur, host, t.shape, t.device)
        if t.is_cuda:
            self.d2h.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.d2h):
                host.copy_(t.reshape(-1), non_blocking=True)
                tk.event = torch.cuda.Event()
                tk.event.record(self.d2h)
            t.record_stream(self.d2h)      # the allocator may reuse the block only after the copy has drained
        else:
            host.copy_(t.detach().reshape(-1))
        self._cur.tickets.append(tk)
        s = self._stats
        s["tensors_offloaded"] += 1
        s["bytes_offloaded"] += t.numel() * t.element_size()
        s["live"] += 1
        return tk

    def _reload(self, tk: _Ticket, stream=None):
        if tk.dev is not None:
            return
        if tk.device.type == "cuda":
            s = stream or torch.cuda.current_stream()
            if tk.event is not None:
                s.wait_event(tk.event)
            with torch.cuda.stream(s):
                tk.dev = torch.empty(tk.shape, dtype=tk.host.dtype, device=tk.device)
                tk.dev.view(-1).copy_(tk.host, non_blocking=True)
        else:
            tk.dev = tk.host.clone().view(tk.shape)

    def _unpack(self, obj):
        if not isinstance(obj, _Ticket):
            return obj
        if obj.dev is None:
            self._stats["sync_reloads"] += 1
            self._reload(obj)
        elif obj.device.type == "cuda" and obj.group.prefetched:
            torch.cuda.current_stream().wait_stream(self.h2d)
        out = obj.dev
        if out.is_cuda:
            out.record_stream(torch.cuda.current_stream())
        obj.uses += 1
        if obj.uses == 1:
            self.pool.put(obj.host)
            self._stats["live"] -= 1
        return out      # the ticket keeps ``dev`` until autograd frees the node that holds it (right after that node's backward)

    # ---- public API ----
    @contextmanager
    def group(self, name: str):
        g = _Group(name, len(self.groups))
        self.groups.append(g)
        self._stats["groups"] += 1
        prev, self._cur = self._cur, g
        try:
            with torch.autograd.graph.saved_tensors_hooks(self._pack, self._unpack):
                yield g
        finally:
            self._cur = prev

    def commit(self, out: torch.Tensor, name: Optional[str] = None) -> torch.Tensor:
        """Close the most recent group; returns ``out`` routed through the node that triggers the prefetch chain in backward."""
        idx = len(self.groups) - 1
        if not out.requires_grad:
            return out
        return _Commit.apply(out, self, idx)

    def _on_backward_reaches(self, index: int) -> None:
        """Backward is about to enter group ``index``: make sure it is resident, and start fetching group ``index-1``."""
        for j in (index, index - 1):
            if j < 0 or self.groups[j].prefetched:
                continue
            g = self.groups[j]
            inflight = 0
            for tk in g.tickets:
                if tk.uses == 0 and tk.dev is None:
                    inflight += tk.host.numel() * tk.host.element_size()
                    if inflight > self.max_inflight:
                        break
                    self._reload(tk, self.h2d)
            g.prefetched = True

    def reset(self) -> None:
        """Call between steps: forget the group records (host buffers stay pooled)."""
        self.groups.clear()
        self._cur = None

    def stats(self) -> dict:
        return dict(self._stats, pool_bytes=self.pool.allocated_bytes)


# functional aliases in the reference's vocabulary
def fine_grained_offloading_group_start(mgr: FineGrainedActivationOffloadingInterface, name: str):
    return mgr.group(name)


def fine_grained_offloading_group_commit(mgr: FineGrainedActivationOffloadingInterface, tensor: torch.Tensor, name: Optional[str] = None):
    return mgr.commit(tensor, name)
