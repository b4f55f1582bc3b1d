"""Pipeline helpers and the fine-grained schedule-node machinery (reference ``pipeline_parallel/utils.py``:
rank helpers :60-121, ``ScheduleNode`` :181, ``AbstractSchedulePlan`` :355, stream registry :381-405).

A *schedule node* is a slice of a layer (attention, MoE dispatch, experts, combine …) that owns its slice of the
autograd graph: inputs are detached on entry, so a node's backward can be run on its own, on its own stream, at
a time the schedule chooses.  That is what lets the forward of one micro-batch overlap the backward of another
(``combined_1f1b.py``): communication nodes live on the communication stream, compute nodes on the compute
stream, and CUDA events express the true data dependencies between consecutive nodes of one micro-batch.

On CPU (tests) streams are ``None`` and every node runs inline; the numerics are identical by construction.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------- rank helpers
def is_pp_first_stage(pp_group) -> bool:
    return pp_group is None or dist.get_rank(pp_group) == 0


def is_pp_last_stage(pp_group) -> bool:
    return pp_group is None or dist.get_rank(pp_group) == dist.get_world_size(pp_group) - 1


def is_vp_first_stage(vp_stage: Optional[int], vp_size: Optional[int]) -> bool:
    return vp_size is None or vp_size <= 1 or vp_stage == 0


def is_vp_last_stage(vp_stage: Optional[int], vp_size: Optional[int]) -> bool:
    return vp_size is None or vp_size <= 1 or vp_stage == vp_size - 1


def _global(pp_group, idx: int) -> int:
    return dist.get_process_group_ranks(pp_group)[idx]


def get_pp_first_rank(pp_group) -> int:
    return _global(pp_group, 0)


def get_pp_last_rank(pp_group) -> int:
    return _global(pp_group, dist.get_world_size(pp_group) - 1)


def get_pp_next_rank(pp_group) -> int:
    return _global(pp_group, (dist.get_rank(pp_group) + 1) % dist.get_world_size(pp_group))


def get_pp_prev_rank(pp_group) -> int:
    return _global(pp_group, (dist.get_rank(pp_group) - 1) % dist.get_world_size(pp_group))


def make_viewless(t):
    """A tensor that does not keep its base alive (so the base's storage can be released early)."""
    if not isinstance(t, torch.Tensor) or t._base is None:
        return t
    out = torch.empty((1,), dtype=t.dtype, device=t.device, requires_grad=t.requires_grad)
    out.data = t.data
    return out


# ---------------------------------------------------------------------------- streams
_STREAMS = {"comp": None, "comm": None}


def set_streams(comp_stream=None, comm_stream=None, high_priority_comm: bool = True) -> None:
    """Register the two streams of the overlapped schedule.  Compute stays on the current stream unless told otherwise;
    the communication stream is created with high priority so short a2a kernels are not queued behind long GEMMs."""
    if not torch.cuda.is_available():
        _STREAMS["comp"] = _STREAMS["comm"] = None
        return
    _STREAMS["comp"] = comp_stream or torch.cuda.current_stream()
    _STREAMS["comm"] = comm_stream or torch.cuda.Stream(priority=-1 if high_priority_comm else 0)


def get_comp_stream():
    return _STREAMS["comp"]


def get_comm_stream():
    return _STREAMS["comm"]


@contextmanager
def _on(stream):
    if stream is None:
        yield
    else:
        with torch.cuda.stream(stream):
            yield


# ---------------------------------------------------------------------------- nodes
class NoopScheduleNode:
    """Placeholder with the node interface (a dense layer has no dispatch / combine)."""

    name = "noop"

    def forward(self, inputs=()):
        return inputs

    def backward(self, grads=()):
        return grads


class ScheduleNode:
    """One independently schedulable slice of the autograd graph.

    ``forward(inputs)`` detaches the inputs (keeping ``requires_grad``), runs ``fn`` on the node's stream and keeps inputs
    and outputs; ``backward(output_grads)`` runs autograd for this slice only and returns the input gradients, releasing
    everything it held.  ``event`` (shared by the nodes of one micro-batch) orders consecutive nodes across streams."""

    def __init__(self, fn: Callable, stream: Optional[Callable] = None, event=None, backward_fn: Optional[Callable] = None,
                 free_input: bool = False, name: str = "node"):
        self.fn, self.backward_fn, self.name, self.free_input = fn, backward_fn, name, free_input
        self._stream, self.event = stream, event
        self.inputs: Optional[List] = None
        self.outputs = None

    @property
    def stream(self):
        return self._stream() if callable(self._stream) else self._stream

    @contextmanager
    def _ordered(self):
        s = self.stream
        if s is None or self.event is None:
            with _on(s):
                yield
            return
        self.event.wait(s)          # everything the previous node of this micro-batch recorded
        with _on(s):
            yield
        self.event.record(s)

    def forward(self, inputs=()):
        if not isinstance(inputs, tuple):
            inputs = (inputs,)
        with self._ordered():
            held = []
            for x in inputs:
                if isinstance(x, torch.Tensor):
                    d = make_viewless(x).detach()
                    d.requires_grad_(x.requires_grad and x.is_floating_point())
                    held.append(d)
                else:
                    held.append(x)
            out = self.fn(*held)
            self.inputs = held
            self.outputs = out if isinstance(out, tuple) else (out,)
            if self.free_input:
                for x in inputs:
                    if isinstance(x, torch.Tensor) and self.stream is not None:
                        x.record_stream(self.stream)
        return out

    def backward(self, output_grads=()):
        if not isinstance(output_grads, tuple):
            output_grads = (output_grads,)
        with self._ordered():
            if self.backward_fn is not None:
                grads = self.backward_fn(self.outputs, output_grads)
            else:
                pairs = [(o, g) for o, g in zip(self.outputs, output_grads) if isinstance(o, torch.Tensor) and o.requires_grad and g is not None]
                if pairs:
                    torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
                grads = tuple(x.grad if isinstance(x, torch.Tensor) else None for x in self.inputs)
            self.inputs = self.outputs = None
        return grads if len(grads) != 1 else grads[0]


class AbstractSchedulePlan(ABC):
    """A model chunk's forward/backward expressed as nodes; ``run`` executes (and overlaps) a forward plan and a backward plan."""

    @classmethod
    @abstractmethod
    def run(cls, f_plan, b_plan, grad=None, **kwargs):
        ...


class StageDispatchBwdGrad(torch.autograd.Function):
    """Identity in forward; in backward hands the gradient to a callback instead of propagating it (used to cut the graph
    between pipeline stages living in one process)."""

    @staticmethod
    def forward(ctx, x, sink):
        ctx.sink = sink
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.sink(g)
        return None, None


__all__ = [n for n in dir() if not n.startswith("_")]
