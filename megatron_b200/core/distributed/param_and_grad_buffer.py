"""Contiguous parameter / gradient buffers with bucketed gradient reduction.

Parity: reference ``distributed/param_and_grad_buffer.py`` (``_ParamAndGradBucket`` :92,
``_ParamAndGradBucketGroup`` :179, ``_ParamAndGradBuffer`` :1005, ``partition_buckets`` :1658).

Design differences:
* The pre-scale (``grad *= 1/dp``), the reduce-scatter and the dtype cast are ONE step
  (``_reduce_bucket``): on B200 the NVLink kernel ``multimem.ld_reduce``-pulls the
  bucket shard from all peers with fp32 accumulation, scales and writes the shard
  (``parallel.nvlink.NVLinkBackend.reduce_scatter_scaled``); with ``torch.distributed``
  it is ``ReduceOp.AVG``/pre-multiplied SUM.
* Buckets are laid out in *reverse registration order* so that they complete in the
  order autograd produces gradients; each bucket is padded to ``lcm(dp, 128)`` elements
  so shards are 256-byte aligned for vectorised NVLink access.
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..utils import get_pg_rank, get_pg_size
from .distributed_data_parallel_config import DistributedDataParallelConfig


class BufferType(Enum):
    PARAM = 1
    GRAD = 2


def shard_buffer(buffer: torch.Tensor, dp_world_size: int) -> List[torch.Tensor]:
    assert buffer.numel() % dp_world_size == 0
    n = buffer.numel() // dp_world_size
    return [buffer[r * n : (r + 1) * n] for r in range(dp_world_size)]


class _ParamAndGradBucket:
    """A contiguous slice of the buffer that is reduced as one unit."""

    def __init__(self, params: List[torch.nn.Parameter], param_data: Optional[torch.Tensor], grad_data: torch.Tensor,
                 offset: int, numel_unpadded: int, gradient_scaling_factor: float, bucket_id: int):
        self.params_list = params
        self.params = set(params)
        self.param_data, self.grad_data = param_data, grad_data
        self.offset, self.numel_unpadded = offset, numel_unpadded
        self.gradient_scaling_factor = gradient_scaling_factor
        self.bucket_id = bucket_id
        self.param_to_index: Dict[torch.nn.Parameter, Tuple[int, int]] = {}


class _ParamAndGradBucketGroup:
    """Buckets whose collectives are issued together; tracks readiness of member params."""

    def __init__(self, buckets: List[_ParamAndGradBucket], ddp_config: DistributedDataParallelConfig, collective_group, collective_group_size: int):
        self.buckets = buckets
        self.ddp_config = ddp_config
        self.group = collective_group
        self.group_size = collective_group_size
        self.rank = get_pg_rank(collective_group) if collective_group is not None else 0
        self.params = set()
        self.param_to_bucket = {}
        for b in buckets:
            for p in b.params:
                self.params.add(p)
                self.param_to_bucket[p] = b
        self.next_param_gather_bucket_group: Optional["_ParamAndGradBucketGroup"] = None
        self.reset()
        self.param_gather_handle = None
        self.param_gather_dispatched = False
        self.grad_reduce_handle = None
        self.is_first_batch = True

    def reset(self):
        self.params_with_grad = set()
        self.is_last_microbatch = True

    # ---- gradient path -----------------------------------------------------------
    def _check_grads(self):
        if self.ddp_config.check_for_nan_in_grad or self.ddp_config.check_for_large_grads:
            for b in self.buckets:
                norm = b.grad_data.float().norm()
                if self.ddp_config.check_for_nan_in_grad and not torch.isfinite(norm):
                    raise RuntimeError(f"found NaN/Inf in local grad norm of bucket #{b.bucket_id} before the data-parallel reduction")
                if self.ddp_config.check_for_large_grads and float(norm) > 1e10:
                    raise RuntimeError(f"unexpectedly large grad norm {float(norm):.3e} in bucket #{b.bucket_id}")

    def _reduce_bucket(self, b: _ParamAndGradBucket, async_op: bool):
        """scale + reduce(-scatter) in one step; returns a handle or None."""
        cfg = self.ddp_config
        nvl = None
        if b.grad_data.is_cuda:
            from ...parallel import collectives

            nvl = collectives.backend_for(self.group)
        # Reference semantics (param_and_grad_buffer.py:600-700): always pre-multiply by gradient_scaling_factor (expert /
        # GTP factors survive average_in_collective); averaging adds a further 1/group_size, fused into the kernel or ReduceOp.AVG.
        scale = b.gradient_scaling_factor
        avg = bool(cfg.average_in_collective)
        if nvl is not None:
            eff = scale / self.group_size if avg else scale
            if cfg.use_distributed_optimizer:
                return nvl.reduce_scatter_scaled_(b.grad_data, eff, async_op=async_op)
            return nvl.all_reduce_scaled_(b.grad_data, eff, async_op=async_op)
        op = dist.ReduceOp.SUM
        if avg:
            if b.grad_data.is_cuda and dist.get_backend(self.group) == "nccl":
                op = dist.ReduceOp.AVG
            else:
                scale = scale / self.group_size          # gloo has no AVG: SUM x 1/group_size
        if scale != 1.0:
            b.grad_data.mul_(scale)
        if cfg.use_distributed_optimizer:
            local = shard_buffer(b.grad_data, self.group_size)[self.rank]
            return dist.reduce_scatter_tensor(local, b.grad_data, op=op, group=self.group, async_op=async_op)
        return dist.all_reduce(b.grad_data, op=op, group=self.group, async_op=async_op)

    def start_grad_sync(self):
        assert self.grad_reduce_handle is None, "should not have multiple communication calls outstanding at once"
        self._check_grads()
        if self.group_size == 1:
            for b in self.buckets:
                if b.gradient_scaling_factor != 1.0:
                    b.grad_data.mul_(b.gradient_scaling_factor)
            self.grad_reduce_handle = None
            return
        async_op = self.ddp_config.overlap_grad_reduce
        handles = [self._reduce_bucket(b, async_op) for b in self.buckets]
        self.grad_reduce_handle = [h for h in handles if h is not None] if async_op else None

    def finish_grad_sync(self):
        self.param_gather_dispatched = False
        if not self.ddp_config.overlap_grad_reduce:
            self.start_grad_sync()
            return
        if self.is_first_batch and self.grad_reduce_handle is None and self.group_size > 1:
            # nothing was launched by hooks (e.g. frozen params): reduce now
            self.start_grad_sync()
        if self.grad_reduce_handle is not None:
            for h in self.grad_reduce_handle:
                h.wait()
            self.grad_reduce_handle = None
        self.is_first_batch = False

    def register_grad_ready(self, param: torch.nn.Parameter):
        assert self.ddp_config.overlap_grad_reduce, "register_grad_ready() is only for overlap_grad_reduce=True"
        if not self.is_last_microbatch:
            return
        assert param in self.param_to_bucket, "param is not in this bucket group"
        assert param not in self.params_with_grad, "cannot set grad twice"
        self.params_with_grad.add(param)
        if len(self.params_with_grad) == len(self.params):
            self.start_grad_sync()

    # ---- parameter path (distributed optimizer) ------------------------------------
    def start_param_sync(self, force_sync: bool = False):
        assert self.ddp_config.use_distributed_optimizer
        if force_sync:
            if self.param_gather_handle is not None:
                for h in self.param_gather_handle:
                    h.wait()
                self.param_gather_handle = None
                return
        else:
            assert self.param_gather_handle is None
        async_op = self.ddp_config.overlap_param_gather and not force_sync
        handles = []
        if self.group_size > 1:
            for b in self.buckets:
                nvl = None
                if b.param_data.is_cuda:
                    from ...parallel import collectives

                    nvl = collectives.backend_for(self.group)
                if nvl is not None:
                    h = nvl.all_gather_inplace_(b.param_data, async_op=async_op)
                else:
                    local = shard_buffer(b.param_data, self.group_size)[self.rank]
                    h = dist.all_gather_into_tensor(b.param_data, local, group=self.group, async_op=async_op)
                if h is not None:
                    handles.append(h)
        self.param_gather_handle = handles if async_op else None
        self.param_gather_dispatched = True

    def finish_param_sync(self, skip_next_bucket_dispatch: bool = False):
        assert self.ddp_config.use_distributed_optimizer and self.ddp_config.overlap_param_gather
        if not self.param_gather_dispatched:
            self.start_param_sync()
        if self.param_gather_handle is not None:
            for h in self.param_gather_handle:
                h.wait()
            self.param_gather_handle = None
            # chain: kick the next bucket group's gather so it overlaps this group's compute
            nxt = self.next_param_gather_bucket_group
            if nxt is not None and not skip_next_bucket_dispatch and not nxt.param_gather_dispatched:
                nxt.start_param_sync()


class _ParamAndGradBuffer:
    """One flat (param, grad) buffer pair for params of equal (param_dtype, grad_dtype)."""

    def __init__(self, ddp_config: DistributedDataParallelConfig, param_dtype: torch.dtype, grad_dtype: torch.dtype,
                 params: List[torch.nn.Parameter], data_parallel_group, bucket_size: Optional[int], param_to_name: Dict,
                 gradient_scaling_factor: float, param_indices: Optional[List[int]] = None):
        self.ddp_config = ddp_config
        self.params = params
        self.param_dtype, self.grad_dtype = param_dtype, grad_dtype
        self.data_parallel_group = data_parallel_group
        self.data_parallel_world_size = get_pg_size(data_parallel_group)
        self.gradient_scaling_factor = gradient_scaling_factor
        self.param_to_name = param_to_name
        assert len(set(params)) == len(params), "duplicate parameter in buffer"
        dp = self.data_parallel_world_size
        self.bucket_align = math.lcm(dp, 128) if ddp_config.use_distributed_optimizer else 64

        def pad(n, mult):
            return (n + mult - 1) // mult * mult

        # ---- layout: iterate params in reverse (≈ order grads become ready) --------
        self.param_index_map: Dict[torch.nn.Parameter, Tuple[int, int, int]] = {}
        self.bucket_indices: List[Tuple[int, int]] = []
        per_bucket_numel_unpadded: List[int] = []
        bucket_params: List[List[torch.nn.Parameter]] = [[]]
        offset = 0
        bucket_start = 0
        bucket_id = 0

        def close_bucket(end_unpadded):
            nonlocal offset, bucket_start, bucket_id
            per_bucket_numel_unpadded.append(end_unpadded - bucket_start)
            end = pad(end_unpadded, self.bucket_align)
            self.bucket_indices.append((bucket_start, end))
            bucket_start = end
            bucket_id += 1
            bucket_params.append([])
            return end

        for p in reversed(params):
            if not p.requires_grad:
                continue
            start = pad(offset, 64) if ddp_config.use_distributed_optimizer else offset  # 128B-aligned params
            # shared embedding gets its own bucket so its (late) grad does not hold up others
            if getattr(p, "shared_embedding", False) and bucket_params[-1]:
                start = close_bucket(offset)
            end = start + p.numel()
            self.param_index_map[p] = (start, end, bucket_id)
            bucket_params[-1].append(p)
            offset = end
            if (bucket_size is not None and (offset - bucket_start) >= bucket_size) or getattr(p, "shared_embedding", False):
                offset = close_bucket(offset)
        if bucket_params[-1]:
            offset = close_bucket(offset)
        else:
            bucket_params.pop()
        self.numel = offset
        self.numel_unpadded = sum(per_bucket_numel_unpadded)
        assert self.numel % dp == 0 or not ddp_config.use_distributed_optimizer

        dev = params[0].device if params else "cpu"
        self.param_data = None
        if ddp_config.use_distributed_optimizer:
            self.param_data = self._alloc(self.numel, param_dtype, dev, "param")
        self.grad_data = self._alloc(self.numel, grad_dtype, dev, "grad")
        self.grad_data.zero_()

        # ---- remap params / create main_grad views ------------------------------------
        self.buckets: List[_ParamAndGradBucket] = []
        for bid, (bs, be) in enumerate(self.bucket_indices):
            b = _ParamAndGradBucket(
                bucket_params[bid], self.param_data[bs:be] if self.param_data is not None else None, self.grad_data[bs:be],
                bs, per_bucket_numel_unpadded[bid], gradient_scaling_factor, bid,
            )
            self.buckets.append(b)
        self.param_to_bucket: Dict[torch.nn.Parameter, _ParamAndGradBucket] = {}
        for p, (s, e, bid) in self.param_index_map.items():
            if self.param_data is not None:
                new = self.param_data[s:e].view(p.shape)
                new.copy_(p.data)
                p.data = new
            p.main_grad = self.grad_data[s:e].view(p.shape)
            b = self.buckets[bid]
            b.param_to_index[p] = (s - b.offset, e - b.offset)
            self.param_to_bucket[p] = b

    def _alloc(self, numel, dtype, device, kind):
        """Buffers that peers must address come from the symmetric heap when one exists."""
        if torch.device(device).type == "cuda" and self.data_parallel_world_size > 1:
            from ...parallel import collectives

            be = collectives.backend_for(self.data_parallel_group)
            if be is not None:
                return be.alloc_symmetric(numel, dtype)
        return torch.zeros(numel, dtype=dtype, device=device)

    def scale_gradients(self, scaling_factor: float) -> None:
        self.grad_data.mul_(scaling_factor)

    def reset(self):
        self.grad_data.zero_()


def partition_buckets(buffers: List[_ParamAndGradBuffer], force_single_bucket_group: bool = False) -> List[_ParamAndGradBucketGroup]:
    """Group buckets for joint communication: one group per bucket by default; a single
    group when bucketing is disabled (reference :1658)."""
    if not buffers:
        return []
    groups = []
    if force_single_bucket_group:
        allb = [b for buf in buffers for b in buf.buckets]
        return [_ParamAndGradBucketGroup(allb, buffers[0].ddp_config, buffers[0].data_parallel_group, buffers[0].data_parallel_world_size)]
    for buf in buffers:
        for b in buf.buckets:
            groups.append(_ParamAndGradBucketGroup([b], buf.ddp_config, buf.data_parallel_group, buf.data_parallel_world_size))
    return groups
