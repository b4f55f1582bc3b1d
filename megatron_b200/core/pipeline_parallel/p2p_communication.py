"""Pipeline-stage point-to-point exchange over NCCL/Gloo p2p
(reference ``pipeline_parallel/p2p_communication.py:145-675``).

Pipeline send/recv intentionally stays on the library p2p path (BASELINE.json); what is
different here is that every exchange is expressed as ONE ``exchange()`` call that posts
all sends before any wait — deadlock-free for any combination of directions — and returns
handles so interleaved 1F1B can defer waits (``overlap_p2p_comm``).
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.distributed as dist

from .. import parallel_state as ps
from ..model_parallel_config import ModelParallelConfig

Shape = Union[List[int], torch.Size, Tuple[int, ...]]


class _Handles:
    def __init__(self, reqs):
        self.reqs = reqs

    def wait(self):
        for r in self.reqs:
            r.wait()
        self.reqs = []


class P2PCommunicator:
    def __init__(self, pp_group=None, config: ModelParallelConfig = None):
        self.pp_group = pp_group if pp_group is not None else ps.get_pipeline_model_parallel_group()
        self.config = config
        ranks = dist.get_process_group_ranks(self.pp_group)
        me = dist.get_rank()
        i = ranks.index(me)
        self.rank_in_group, self.size = i, len(ranks)
        self.next_rank = ranks[(i + 1) % len(ranks)]
        self.prev_rank = ranks[(i - 1) % len(ranks)]
        self.is_first = i == 0
        self.is_last = i == len(ranks) - 1

    # ---- core ------------------------------------------------------------------------
    def _device(self):
        if dist.get_backend(self.pp_group) == "gloo" or not torch.cuda.is_available():
            return torch.device("cpu")
        return torch.device("cuda", torch.cuda.current_device())

    def _exchange_shapes(self, send_next, send_prev, recv_prev: bool, recv_next: bool):
        """variable_seq_lengths: hand-shake 3-int shapes first."""
        dev = self._device()

        def enc(t):
            return torch.tensor(list(t.shape) + [0] * (3 - t.dim()), dtype=torch.int64, device=dev) if t is not None else None

        sp = torch.empty(3, dtype=torch.int64, device=dev) if recv_prev else None
        sn = torch.empty(3, dtype=torch.int64, device=dev) if recv_next else None
        h = self._post(enc(send_next), enc(send_prev), sp, sn)
        h.wait()
        return (sp.tolist() if sp is not None else None), (sn.tolist() if sn is not None else None)

    def _post(self, send_next, send_prev, recv_prev_buf, recv_next_buf) -> _Handles:
        """Post every send/recv; even stages send first and odd stages receive first so that
        blocking back ends (gloo) pair up; NCCL gets one batched group."""
        ops = []
        g = self.pp_group
        use_batch = self.config is None or self.config.batch_p2p_comm
        if dist.get_backend(g) != "gloo" and use_batch:
            # per-peer matching is positional inside an NCCL group: with pp == 2 both
            # directions go to the same peer, so sends are ordered (next, prev) and
            # receives (prev, next) on every rank.
            if send_next is not None:
                ops.append(dist.P2POp(dist.isend, send_next, self.next_rank, g))
            if send_prev is not None:
                ops.append(dist.P2POp(dist.isend, send_prev, self.prev_rank, g))
            if recv_prev_buf is not None:
                ops.append(dist.P2POp(dist.irecv, recv_prev_buf, self.prev_rank, g))
            if recv_next_buf is not None:
                ops.append(dist.P2POp(dist.irecv, recv_next_buf, self.next_rank, g))
            return _Handles(dist.batch_isend_irecv(ops) if ops else [])
        reqs = []
        even = self.rank_in_group % 2 == 0
        order = [("sn", send_next), ("rp", recv_prev_buf), ("sp", send_prev), ("rn", recv_next_buf)] if even else \
                [("rp", recv_prev_buf), ("sn", send_next), ("rn", recv_next_buf), ("sp", send_prev)]
        for kind, t in order:
            if t is None:
                continue
            if kind == "sn":
                reqs.append(dist.isend(t, self.next_rank, group=g, tag=0))
            elif kind == "sp":
                reqs.append(dist.isend(t, self.prev_rank, group=g, tag=1))
            elif kind == "rp":
                reqs.append(dist.irecv(t, self.prev_rank, group=g, tag=0))
            else:
                reqs.append(dist.irecv(t, self.next_rank, group=g, tag=1))
        return _Handles(reqs)

    def exchange(self, *, send_next=None, send_prev=None, recv_prev: bool = False, recv_next: bool = False,
                 tensor_shape: Shape = None, wait: bool = True):
        """Returns ``(from_prev, from_next, handles)``; tensors are usable after ``handles.wait()``."""
        cfg = self.config
        dev = self._device()
        dtype = (cfg.pipeline_dtype if cfg is not None and cfg.pipeline_dtype is not None else torch.float32)
        prev_shape = next_shape = tensor_shape
        if cfg is not None and cfg.variable_seq_lengths:
            ps_, ns_ = self._exchange_shapes(send_next, send_prev, recv_prev, recv_next)
            prev_shape = ps_ if ps_ is not None else prev_shape
            next_shape = ns_ if ns_ is not None else next_shape
        from_prev = torch.empty(tuple(prev_shape), dtype=dtype, device=dev, requires_grad=True) if recv_prev else None
        from_next = torch.empty(tuple(next_shape), dtype=dtype, device=dev, requires_grad=True) if recv_next else None
        sn = send_next.contiguous().to(dev) if send_next is not None else None
        sp = send_prev.contiguous().to(dev) if send_prev is not None else None
        h = self._post(sn, sp, from_prev.detach() if from_prev is not None else None, from_next.detach() if from_next is not None else None)
        if wait:
            h.wait()
            if dev.type == "cuda" and cfg is not None and cfg.batch_p2p_comm and cfg.batch_p2p_sync:
                torch.cuda.synchronize()
            return from_prev, from_next, None
        return from_prev, from_next, h

    # ---- named wrappers (reference API) ----------------------------------------------------
    def recv_forward(self, tensor_shape, is_first_stage: Optional[bool] = None):
        if self.is_first if is_first_stage is None else is_first_stage:
            return None
        return self.exchange(recv_prev=True, tensor_shape=tensor_shape)[0]

    def recv_backward(self, tensor_shape, is_last_stage: Optional[bool] = None):
        if self.is_last if is_last_stage is None else is_last_stage:
            return None
        return self.exchange(recv_next=True, tensor_shape=tensor_shape)[1]

    def send_forward(self, output_tensor, is_last_stage: Optional[bool] = None):
        if not (self.is_last if is_last_stage is None else is_last_stage):
            self.exchange(send_next=output_tensor)

    def send_backward(self, input_tensor_grad, is_first_stage: Optional[bool] = None):
        if not (self.is_first if is_first_stage is None else is_first_stage):
            self.exchange(send_prev=input_tensor_grad)

    def send_forward_recv_backward(self, output_tensor, tensor_shape, is_last_stage: Optional[bool] = None):
        if self.is_last if is_last_stage is None else is_last_stage:
            return None
        return self.exchange(send_next=output_tensor, recv_next=True, tensor_shape=tensor_shape)[1]

    def send_backward_recv_forward(self, input_tensor_grad, tensor_shape, is_first_stage: Optional[bool] = None):
        if self.is_first if is_first_stage is None else is_first_stage:
            return None
        return self.exchange(send_prev=input_tensor_grad, recv_prev=True, tensor_shape=tensor_shape)[0]

    def send_forward_recv_forward(self, output_tensor, recv_prev: bool, tensor_shape, overlap_p2p_comm: bool = False):
        fp, _, h = self.exchange(send_next=output_tensor, recv_prev=recv_prev, tensor_shape=tensor_shape, wait=not overlap_p2p_comm)
        return (fp, h) if overlap_p2p_comm else fp

    def send_backward_recv_backward(self, input_tensor_grad, recv_next: bool, tensor_shape, overlap_p2p_comm: bool = False):
        _, fn, h = self.exchange(send_prev=input_tensor_grad, recv_next=recv_next, tensor_shape=tensor_shape, wait=not overlap_p2p_comm)
        return (fn, h) if overlap_p2p_comm else fn

    def send_forward_backward_recv_forward_backward(self, output_tensor, input_tensor_grad, recv_prev: bool, recv_next: bool, tensor_shape):
        fp, fn, _ = self.exchange(send_next=output_tensor, send_prev=input_tensor_grad, recv_prev=recv_prev, recv_next=recv_next, tensor_shape=tensor_shape)
        return fp, fn
