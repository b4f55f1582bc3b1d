"""``MegatronFSDP``: the module wrapper that drives ``ParamAndGradBuffer`` through forward / backward
(reference ``distributed/fsdp/src/megatron_fsdp/megatron_fsdp.py:1-1561``).

Sharding strategies (``--data-parallel-sharding-strategy``):

==================== ========= ========= =============== ===========================================================
strategy             weights   gradients optimizer state communication per step
==================== ========= ========= =============== ===========================================================
``no_shard``         resident  resident  replicated      all-reduce(grads)
``optim``            resident  resident  sharded         reduce-scatter(grads) + all-gather(weights)          ZeRO-1
``optim_grads``      resident  sharded   sharded         per-bucket reduce-scatter in backward + all-gather   ZeRO-2
``optim_grads_params`` sharded sharded   sharded         + per-unit all-gather in forward AND backward        ZeRO-3
==================== ========= ========= =============== ===========================================================

Hook protocol for ZeRO-3 (one FSDP unit = one module of ``fsdp_unit_modules``, usually a transformer layer):
forward-pre: gather the unit (+ prefetch the following units up to ``suggested_AG_prefetch_size`` elements), wait, point the
parameters at the bucket; forward-post: release.  backward-pre: gather again, prefetching in REVERSE order; a parameter's
post-accumulate-grad hook hands its gradient to the ``GradReducePipeline``; when the unit's last gradient arrived its weight
bucket is released.  Parameters outside any unit (embeddings, final norm, head) are gathered at the root and released at the
end of the backward pass."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .param_and_grad_buffer import AllGatherPipeline, BucketingPolicy, BucketStatus, GradReducePipeline, ParamAndGradBuffer, PrefetchOrder


class MegatronFSDP(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, ddp_config=None, fsdp_unit_modules: Optional[Sequence[type]] = None,
                 data_parallel_sharding_strategy: Optional[str] = None, dp_group=None, expert_dp_group=None, outer_dp_group=None,
                 preserve_fp32_weights: bool = True, grad_reduce_in_fp32: bool = True, overlap_param_gather: bool = True, overlap_grad_reduce: bool = True,
                 suggested_AG_prefetch_size: Optional[int] = None, suggested_RS_queue_capacity: Optional[int] = None, allocator: str = "auto",
                 double_buffer_size: int = 2, average_gradients: bool = True, device=None):
        super().__init__()
        self.module, self.ddp_config = module, ddp_config
        strategy = data_parallel_sharding_strategy or getattr(ddp_config, "data_parallel_sharding_strategy", None) or "optim_grads_params"
        self.strategy = strategy
        self.dp_group = dp_group
        self.outer_dp_group = outer_dp_group                 # HSDP: shard inside dp_group (one NVLink domain), replicate across this one
        world = dist.get_world_size(dp_group) if dist.is_initialized() else 1
        outer = dist.get_world_size(outer_dp_group) if (outer_dp_group is not None and dist.is_initialized()) else 1
        scale = 1.0 / (world * outer) if average_gradients else None
        unit_types = tuple(fsdp_unit_modules) if fsdp_unit_modules else ()
        policy = BucketingPolicy(suggested_bucket_size=getattr(ddp_config, "bucket_size", None) or 40_000_000, fsdp_unit_modules=unit_types,
                                 data_parallel_sharding_strategy=strategy)
        self.param_and_grad_buffer = ParamAndGradBuffer(ddp_config, module, policy, dp_group, expert_dp_group, preserve_fp32_weights, grad_reduce_in_fp32,
                                                        scale, scale, device, allocator=allocator, double_buffer_size=double_buffer_size)
        buf = self.param_and_grad_buffer
        self.all_gather_pipeline = AllGatherPipeline(buf, async_op=overlap_param_gather)
        self.grad_reduce_pipeline = GradReducePipeline(buf, check_nans=bool(getattr(ddp_config, "check_for_nan_in_grad", False)),
                                                       suggested_queue_capacity=suggested_RS_queue_capacity)
        self.overlap_grad_reduce = overlap_grad_reduce
        self.shard_weights = strategy == "optim_grads_params"
        self.shard_grads = strategy in ("optim_grads", "optim_grads_params")
        # look-ahead of one unit by default: with a double-buffered pool that is all the pool can hold anyway
        unit_sizes = [sum(buf.parameter_groups[b].model_weight_buffer.bucket_index.size for b in bids) for bids in buf.fsdp_unit_buckets.values()]
        self.suggested_AG_prefetch_size = suggested_AG_prefetch_size if suggested_AG_prefetch_size is not None else (max(unit_sizes) if unit_sizes else 0)
        self.is_last_microbatch = True
        self._unit_params: Dict[torch.nn.Module, List[torch.nn.Parameter]] = {}
        self._root_params: List[torch.nn.Parameter] = []
        self._pending: Dict[int, set] = {}
        self._register_hooks(unit_types)

    # ---- hooks --------------------------------------------------------------------------------------------------------
    def _register_hooks(self, unit_types):
        buf = self.param_and_grad_buffer
        claimed = set()
        for m in self.module.modules():
            if unit_types and isinstance(m, unit_types):
                ps = [p for p in m.parameters() if id(p) not in claimed]
                if ps:
                    claimed.update(id(p) for p in ps)
                    self._unit_params[m] = ps
        self._root_params = [p for p in self.module.parameters() if id(p) not in claimed]
        for m, ps in self._unit_params.items():
            m.register_forward_pre_hook(self._make_pre_forward(ps))
            m.register_forward_hook(self._make_post_forward(ps))
            m.register_full_backward_pre_hook(self._make_pre_backward(ps))
        for gi, g in enumerate(buf.parameter_groups):
            if not g.requires_grad:
                continue
            for p in g.params:
                p.register_post_accumulate_grad_hook(self._make_grad_hook(gi))

    def _buckets_of(self, params) -> List[int]:
        m = self.param_and_grad_buffer.param_to_param_group
        return sorted({m[id(p)] for p in params})

    def _gather_and_wait(self, params, order: PrefetchOrder):
        ag = self.all_gather_pipeline
        ag.all_gather_params(params, prefetch=self.shard_weights, prefetch_order=order, suggested_AG_prefetch_size=self.suggested_AG_prefetch_size)
        for b in self._buckets_of(params):
            ag.wait_bucket_ready(b)

    def _make_pre_forward(self, params):
        def hook(mod, args):
            self._gather_and_wait(params, PrefetchOrder.FORWARD_PASS_ORDER)
        return hook

    def _make_post_forward(self, params):
        def hook(mod, args, out):
            if self.shard_weights:
                for b in self._buckets_of(params):
                    self.all_gather_pipeline.release_bucket(b)
            return out
        return hook

    def _make_pre_backward(self, params):
        def hook(mod, grad_out):
            self._gather_and_wait(params, PrefetchOrder.BACKWARD_PASS_ORDER)
        return hook

    def _make_grad_hook(self, gi: int):
        buf = self.param_and_grad_buffer
        g = buf.parameter_groups[gi]

        def hook(p):
            arrived = self._pending.setdefault(gi, set())
            arrived.add(id(p))
            if len(arrived) < len(g.params):
                return                                       # (a group with a parameter that never gets a gradient is flushed in finish_grad_sync)
            self._pending.pop(gi, None)
            self.grad_reduce_pipeline.reduce_gradients(g.params, async_op=self.overlap_grad_reduce)
            if self.shard_weights and g.fsdp_unit_id is not None:
                self.all_gather_pipeline.release_bucket(gi)
        return hook

    # ---- forward / step protocol ------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        if self._root_params:
            self._gather_and_wait(self._root_params, PrefetchOrder.FORWARD_PASS_ORDER)
        return self.module(*args, **kwargs)

    @contextmanager
    def no_sync(self):
        """Resident-gradient strategies defer their (single) reduction to the last micro-batch; sharded-gradient strategies
        reduce every micro-batch into the fp32 shard accumulator, so there is nothing to defer."""
        prev, self.is_last_microbatch = self.is_last_microbatch, False
        try:
            yield
        finally:
            self.is_last_microbatch = prev

    def start_grad_sync(self, *unused):
        pass

    def finish_grad_sync(self, force_all_reduce: bool = False):
        buf = self.param_and_grad_buffer
        # parameters whose gradient never arrived (unused in this step) still belong to complete buckets
        for gi in list(self._pending):
            self._pending.pop(gi)
            self.grad_reduce_pipeline.reduce_gradients([p for p in buf.parameter_groups[gi].params if p.grad is not None], force=True)
        self.grad_reduce_pipeline.wait_for_previous_grad_reduce(0)
        if not self.shard_grads:
            if self.strategy == "optim" and not force_all_reduce:
                buf.reduce_scatter_gradients()
            else:
                for h in buf.all_reduce_gradients():
                    h.wait()
        if self.outer_dp_group is not None and dist.get_world_size(self.outer_dp_group) > 1:
            for g in buf.parameter_groups:
                if g.main_grad_buffer is not None:
                    t = g.main_grad_buffer.get_shard_from_local_buffer() if self.strategy != "no_shard" else g.main_grad_buffer.data
                    dist.all_reduce(t, group=self.outer_dp_group)
        if self.shard_weights:
            self.all_gather_pipeline.reset()
        buf.update_main_grads()

    def zero_grad_buffer(self):
        self.param_and_grad_buffer.zero_grad()
        self._pending.clear()

    def scale_gradients(self, factor: float):
        self.param_and_grad_buffer.scale_gradients(factor)

    def optimizer_parameters(self) -> List[torch.nn.Parameter]:
        return self.param_and_grad_buffer.optimizer_parameters()

    @torch.no_grad()
    def install_optimized_model_weights(self):
        """After ``optimizer.step()`` on the shard parameters."""
        buf = self.param_and_grad_buffer
        buf.copy_main_weights_to_model_weights()
        if not self.shard_weights:
            # every rank updated only its slice of the resident weights (groups without main weights were updated in full)
            for h in buf.all_gather_parameters(async_op=False):
                h.wait()

    post_optimizer_step = install_optimized_model_weights

    # ---- full state (checkpoint export, tests) ------------------------------------------------------------------------------
    @torch.no_grad()
    def gather_full_state_dict(self) -> Dict[str, torch.Tensor]:
        ag, buf = self.all_gather_pipeline, self.param_and_grad_buffer
        out: Dict[str, torch.Tensor] = {}
        for gi, g in enumerate(buf.parameter_groups):
            was_empty = ag.status[gi] == BucketStatus.EMPTY
            ag.async_bucket_gather(gi)
            ag.wait_bucket_ready(gi)
            for p in g.params:
                out[buf.param_to_name[id(p)]] = p.detach().clone()
            if was_empty and self.shard_weights:
                ag.release_bucket(gi)
        for n, b in self.module.named_buffers():
            out[n] = b.detach().clone()
        return out

    def state_dict(self, *a, **k):
        return self.gather_full_state_dict()

    @torch.no_grad()
    def load_full_state_dict(self, sd: Dict[str, torch.Tensor]):
        buf = self.param_and_grad_buffer
        for g in buf.parameter_groups:
            for i, p in enumerate(g.params):
                full = sd[buf.param_to_name[id(p)]]
                g.model_weight_buffer.set_item(i, full.to(g.dtype))
                if g.main_weight_buffer is not None:
                    g.main_weight_buffer.set_item(i, full.to(g.main_weight_buffer.dtype))
