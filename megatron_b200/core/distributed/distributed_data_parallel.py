"""Data-parallel wrapper over contiguous grad buffers
(reference ``distributed/distributed_data_parallel.py:87-721``)."""
from __future__ import annotations

import logging
from contextlib import contextmanager
from typing import Dict, List

import torch
import torch.distributed as dist

from .. import parallel_state as ps
from ..transformer.transformer_config import TransformerConfig
from ..utils import get_pg_size
from .data_parallel_base import _BaseDataParallel
from .distributed_data_parallel_config import DistributedDataParallelConfig
from .param_and_grad_buffer import _ParamAndGradBuffer, partition_buckets

logger = logging.getLogger(__name__)


class DistributedDataParallel(_BaseDataParallel):
    """Keeps gradients in contiguous fp32/bf16 buffers (``param.main_grad``), accumulates
    into them from autograd hooks, and reduces each bucket over the data-parallel group as
    soon as all its gradients for the last micro-batch are ready."""

    def __init__(self, config: TransformerConfig, ddp_config: DistributedDataParallelConfig, module: torch.nn.Module,
                 disable_bucketing: bool = False, pg_collection=None, full_param_layout=None):
        super().__init__(config=config, module=module)
        self.ddp_config = ddp_config
        if ddp_config.bucket_size is None:
            # bucket must amortise launch latency, not link count: NVSwitch gives full BW per peer
            dp = ps.get_data_parallel_world_size(with_context_parallel=True)
            ddp_config.bucket_size = max(40_000_000, 1_000_000 * dp)
        if not ddp_config.overlap_grad_reduce:
            ddp_config.bucket_size = None
        self.bucket_size = None if disable_bucketing else ddp_config.bucket_size
        if pg_collection is None:
            self.dp_group = ps.get_data_parallel_group(with_context_parallel=True, partial_data_parallel=False) if ps.is_initialized() else None
            self.intra_dp_group = ps.get_data_parallel_group(with_context_parallel=True, partial_data_parallel=True) if ps.is_initialized() else None
            self.expt_dp_group = ps.get_expert_data_parallel_group() if ps.is_initialized() else None
            self.tp_group = ps.get_tensor_model_parallel_group() if ps.is_initialized() else None
            self.pp_group = ps.get_pipeline_model_parallel_group() if ps.is_initialized() else None
            self.ep_group = ps.get_expert_model_parallel_group() if ps.is_initialized() else None
        else:
            self.dp_group = pg_collection.dp_cp
            self.intra_dp_group = getattr(pg_collection, "intra_dp_cp", pg_collection.dp_cp)
            self.expt_dp_group = getattr(pg_collection, "expt_dp", None)
            self.tp_group, self.pp_group, self.ep_group = pg_collection.tp, pg_collection.pp, getattr(pg_collection, "ep", None)

        self.param_to_name: Dict[torch.nn.Parameter, str] = {}
        dense, expert, gtp, egtp = [], [], [], []
        for name, p in self.module.named_parameters():
            if not p.requires_grad:
                continue
            p.grad_added_to_main_grad = False
            self.param_to_name[p] = name
            if getattr(p, "gtp_sharded", False) and getattr(p, "gtp_expert", False):
                egtp.append(p)      # expert-side GTP shard: reduce over the expert replicas of this shard
            elif getattr(p, "gtp_sharded", False):
                gtp.append(p)       # shard of a GTP weight: its gradient is already summed over the remat group, reduce over the orthogonal replicas only
            else:
                (dense if getattr(p, "allreduce", True) else expert).append(p)

        dp_ws = get_pg_size(self.dp_group)
        expt_ws = get_pg_size(self.expt_dp_group)
        if config.calculate_per_token_loss:
            dense_scale = expert_scale = 1.0
        elif ddp_config.average_in_collective:
            dense_scale, expert_scale = 1.0, (expt_ws / dp_ws if dp_ws else 1.0)
        else:
            dense_scale = 1.0 / dp_ws
            expert_scale = 1.0 / dp_ws  # experts see 1/ep of the tokens ep times → same 1/dp

        self.buffers = self._allocate(dense, self.intra_dp_group if ddp_config.num_distributed_optimizer_instances > 1 else self.dp_group, dense_scale)
        self.expert_parallel_buffers = self._allocate(expert, self.expt_dp_group, expert_scale)
        if gtp:
            # buffers with their own reduction group live next to the expert ones (everything downstream iterates both lists)
            gtp_group = ps.get_data_parallel_group_without_gtp()
            gtp_scale = dense_scale if not ddp_config.average_in_collective else get_pg_size(gtp_group) / max(dp_ws, 1)
            self.expert_parallel_buffers += self._allocate(gtp, gtp_group, gtp_scale)
        if egtp:
            eg = ps.get_expert_data_parallel_group_without_gtp()
            e_scale = expert_scale if not ddp_config.average_in_collective else get_pg_size(eg) / max(dp_ws, 1)
            self.expert_parallel_buffers += self._allocate(egtp, eg, e_scale)
        single = self.bucket_size is None
        self.bucket_groups = partition_buckets(self.buffers, force_single_bucket_group=single)
        self.expert_parallel_bucket_groups = partition_buckets(self.expert_parallel_buffers, force_single_bucket_group=single)
        if ddp_config.use_distributed_optimizer and ddp_config.overlap_param_gather:
            for groups in (self.bucket_groups, self.expert_parallel_bucket_groups):
                # forward consumes params in registration order = reverse bucket order
                for i in range(1, len(groups)):
                    groups[len(groups) - i].next_param_gather_bucket_group = groups[len(groups) - i - 1]
        self.param_to_bucket_group = {}
        for g in self.bucket_groups + self.expert_parallel_bucket_groups:
            for p in g.params:
                self.param_to_bucket_group[p] = g

        # ---- hooks ----------------------------------------------------------------
        self.grad_accs = []
        for p in self.module.parameters():
            if p.requires_grad:
                if hasattr(p, "register_post_accumulate_grad_hook"):
                    p.register_post_accumulate_grad_hook(self._make_backward_post_hook(p))
                else:  # pragma: no cover
                    acc = p.expand_as(p).grad_fn.next_functions[0][0]
                    acc.register_hook(lambda *_, _p=p: self._make_backward_post_hook(_p)(_p))
                    self.grad_accs.append(acc)
        self.use_forward_hook = ddp_config.use_distributed_optimizer and ddp_config.overlap_param_gather
        self.remove_forward_pre_hook_handles = {}
        if self.use_forward_hook:
            self.enable_forward_pre_hook()
        self.overlap_param_gather_with_optimizer_step = False

    def _allocate(self, params: List[torch.nn.Parameter], group, scale: float) -> List[_ParamAndGradBuffer]:
        by_dtype: Dict = {}
        for p in params:
            gdt = torch.float32 if self.ddp_config.grad_reduce_in_fp32 else p.dtype
            by_dtype.setdefault((p.dtype, gdt), []).append(p)
        out = []
        for (pdt, gdt), ps_ in by_dtype.items():
            out.append(_ParamAndGradBuffer(self.ddp_config, pdt, gdt, ps_, group, self.bucket_size, self.param_to_name, scale))
        return out

    # ---- hooks ----------------------------------------------------------------------
    def _make_backward_post_hook(self, param: torch.nn.Parameter):
        def hook(*unused):
            if param in self.param_to_bucket_group:
                if param.grad is not None and not param.grad_added_to_main_grad:
                    param.main_grad.add_(param.grad.data)
                param.grad = None
                param.grad_added_to_main_grad = False
                if self.ddp_config.overlap_grad_reduce:
                    self.param_to_bucket_group[param].register_grad_ready(param)

        return hook

    def enable_forward_pre_hook(self):
        assert self.use_forward_hook and not self.remove_forward_pre_hook_handles
        for m in self.module.modules():
            self.remove_forward_pre_hook_handles[m] = m.register_forward_pre_hook(self._make_forward_pre_hook())

    def disable_forward_pre_hook(self, param_sync: bool = True):
        for h in self.remove_forward_pre_hook_handles.values():
            h.remove()
        self.remove_forward_pre_hook_handles = {}
        if param_sync:
            self.start_param_sync(force_sync=True)

    def _make_forward_pre_hook(self):
        def hook(module, *unused):
            for p in module.parameters(recurse=False):
                g = self.param_to_bucket_group.get(p)
                if g is not None and (g.param_gather_handle is not None or (self._param_sync_pending and not g.param_gather_dispatched)):
                    g.finish_param_sync(skip_next_bucket_dispatch=self.ddp_config.align_param_gather)

        return hook

    _param_sync_pending = False

    # ---- public API -----------------------------------------------------------------
    @contextmanager
    def no_sync(self):
        groups = self.bucket_groups + self.expert_parallel_bucket_groups
        for g in groups:
            g.is_last_microbatch = False
        try:
            yield
        finally:
            for g in groups:
                g.is_last_microbatch = True

    def start_param_sync(self, *unused, force_sync: bool = False, force_dispatch: bool = False):
        if not force_sync and self.overlap_param_gather_with_optimizer_step and not force_dispatch:
            return
        self._param_sync_pending = not force_sync
        groups = self.bucket_groups + self.expert_parallel_bucket_groups
        if force_sync or not self.ddp_config.overlap_param_gather:
            for g in groups:
                g.start_param_sync(force_sync=True) if g.param_gather_handle is not None else g.start_param_sync(force_sync=force_sync)
            self._param_sync_pending = False
            return
        # overlapped: only dispatch the first group consumed by forward; the rest chain
        for lst in (self.bucket_groups, self.expert_parallel_bucket_groups):
            if lst:
                lst[-1].start_param_sync()

    def start_grad_sync(self, *unused):
        for g in self.bucket_groups + self.expert_parallel_bucket_groups:
            g.start_grad_sync()

    def finish_grad_sync(self, force_all_reduce: bool = False):
        for g in self.bucket_groups + self.expert_parallel_bucket_groups:
            g.finish_grad_sync()

    def scale_gradients(self, scaling_factor: float):
        for b in self.buffers + self.expert_parallel_buffers:
            b.scale_gradients(scaling_factor)

    def zero_grad_buffer(self):
        for p in self.param_to_name:
            p.grad_added_to_main_grad = False
        for b in self.buffers + self.expert_parallel_buffers:
            b.reset()
        for g in self.bucket_groups + self.expert_parallel_bucket_groups:
            g.reset()

    def broadcast_params(self):
        for p in self.module.parameters():
            group = self.expt_dp_group if not getattr(p, "allreduce", True) else self.dp_group
            if getattr(p, "gtp_sharded", False):
                group = ps.get_expert_data_parallel_group_without_gtp() if getattr(p, "gtp_expert", False) else ps.get_data_parallel_group_without_gtp()
            if group is None or get_pg_size(group) == 1:
                continue
            dist.broadcast(p.data, src=dist.get_process_group_ranks(group)[0], group=group)
