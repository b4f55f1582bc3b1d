"""Run a ``ReshardPlan`` through a copy service (reference ``resharding/execution.py:21-260``)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .copy_services.base import CopyService
from .transforms import ReshardTransform
from .utils import ReshardPlan, get_refit_tensor_dict


def _refresh_module_caches(dst_module: Optional[torch.nn.Module]) -> None:
    """Derived state that depends on the weights (cached transposes, fp8 copies, absorbed MLA matrices) is dropped."""
    if dst_module is None:
        return
    for m in dst_module.modules():
        for hook in ("invalidate_weight_caches", "_invalidate_weight_cache", "reset_weight_cache"):
            fn = getattr(m, hook, None)
            if callable(fn):
                fn()


class _Writeback:
    """A receive that needs work after the wire transfer (transforms): buffers now, ``finalize`` later."""
    __slots__ = ("name", "dst_slice", "buffers")

    def __init__(self, name: str, dst_slice: Tuple[slice, ...], buffers: List[torch.Tensor]):
        self.name, self.dst_slice, self.buffers = name, dst_slice, buffers


@torch.no_grad()
def execute_reshard_plan(plan: ReshardPlan, src_module: Optional[torch.nn.Module], dst_module: Optional[torch.nn.Module], service: CopyService,
                         group=None, transform: Optional[ReshardTransform] = None) -> None:
    """Collective over the service's group.  ``src_module`` / ``dst_module`` may be ``None`` on ranks that only receive / send."""
    transform = transform if transform is not None else plan.transform
    src = get_refit_tensor_dict(src_module) if src_module is not None else {}
    dst = get_refit_tensor_dict(dst_module) if dst_module is not None else {}
    # the peer runs its transform decision on ITS name of the tensor; both sides must agree, so the decision is made on the
    # canonical wire content: a transformed tensor always travels as ONE bf16 piece per op
    for op in plan.send_ops:
        t = src[op.param_name]
        if transform is not None and transform.should_transform(op.param_name):
            pieces = transform.prepare_send(op.param_name, op.my_slice, t)
        else:
            pieces = [t.detach()[op.my_slice]]
        for p in pieces:
            service.submit_send(p, op.peer_rank, op.task_id)
    writebacks: List[_Writeback] = []
    for op in plan.recv_ops:
        if transform is not None and transform.should_transform(op.param_name):
            bufs = transform.prepare_recv(op.param_name, op.my_slice)
            writebacks.append(_Writeback(op.param_name, op.my_slice, bufs))
            for b in bufs:
                service.submit_recv(b, op.peer_rank, op.task_id)
            continue
        view = dst[op.param_name].detach()[op.my_slice]
        service.submit_recv(view, op.peer_rank, op.task_id)
        if op.wire_dtype is not None and op.wire_dtype != view.dtype:
            service.recv_ops[-1].wire_dtype = op.wire_dtype     # the service converts while unpacking
    service.run()
    for wb in writebacks:
        transform.finalize_recv(wb.name, wb.dst_slice, wb.buffers)
    if transform is not None and hasattr(transform, "finish"):
        transform.finish()
    _refresh_module_caches(dst_module)
