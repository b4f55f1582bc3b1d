"""ZeRO-3 fully-sharded data parallel (reference ``distributed/fsdp/src/megatron_fsdp/`` — 13 kLoC; strategy
``optim_grads_params``).

Unit of sharding = one *FSDP unit* (by default every ``TransformerLayer``; everything else forms the root unit).  Per unit:

* ``flat``    — the unit's parameters laid out back-to-back in ONE buffer, padded to a multiple of the DP size.  The model's
  ``nn.Parameter`` objects are views into it.  Outside the unit's compute its storage is released
  (``untyped_storage().resize_(0)``); only this rank's ``shard`` stays resident (+ fp32 master / optimizer state on it).
* forward pre-hook: all-gather ``shard`` → ``flat`` (prefetch of the NEXT unit is issued before compute starts);
  forward post-hook: release ``flat``.
* backward pre-hook: all-gather again; when the last parameter gradient of the unit has been accumulated:
  reduce-scatter the flat gradient (fp32 accumulation into ``main_grad_shard``), release ``flat`` and the full grads.

On B200 the 180 GB HBM rarely *forces* ZeRO-3 for ≤70B models, so the default policy keeps ``keep_fp8_or_bf16_params_resident``
off and sizes the all-gather per layer (~400 MB for Llama-70B / DP=8: ≈0.5 ms on NVLink 5) to overlap with the previous
layer's compute on a side stream.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from ... import parallel_state as ps


class _Unit:
    def __init__(self, name: str, params: List[torch.nn.Parameter], group, main_dtype=torch.float32):
        self.name, self.params, self.group = name, params, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numels = [p.numel() for p in params]
        total = sum(self.numels)
        self.padded = (total + self.world - 1) // self.world * self.world
        self.shard_size = self.padded // self.world
        p0 = params[0]
        self.dtype, self.device = p0.dtype, p0.device
        self.flat = torch.zeros(self.padded, dtype=self.dtype, device=self.device)
        off = 0
        for p, n in zip(params, self.numels):
            self.flat[off : off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off : off + n].view(p.shape)
            off += n
        lo = self.rank * self.shard_size
        # the persistent state: this rank's slice of the parameters (model dtype) + fp32 master + fp32 grad accumulator
        self.shard = self.flat[lo : lo + self.shard_size].clone()
        self.master = torch.nn.Parameter(self.shard.to(main_dtype), requires_grad=True)
        self.main_grad_shard = torch.zeros(self.shard_size, dtype=main_dtype, device=self.device)
        self.resident = True
        self.pending = 0
        self.handle = None

    # ---- parameter residency -----------------------------------------------------------------------------
    def gather(self, async_op: bool = False):
        if self.resident:
            return
        self.flat.untyped_storage().resize_(self.padded * self.flat.element_size())
        self.handle = dist.all_gather_into_tensor(self.flat, self.shard, group=self.group, async_op=async_op)
        self.resident = True

    def wait(self):
        if self.handle is not None:
            self.handle.wait()
            self.handle = None

    def release(self):
        if not self.resident:
            return
        self.wait()
        self.flat.untyped_storage().resize_(0)
        self.resident = False

    # ---- gradients ----------------------------------------------------------------------------------------------
    def reduce_scatter_grads(self, scale: float):
        flat_g = torch.zeros(self.padded, dtype=torch.float32, device=self.device)
        off = 0
        for p, n in zip(self.params, self.numels):
            if p.grad is not None:
                flat_g[off : off + n].copy_(p.grad.reshape(-1))
                p.grad = None
            off += n
        if scale != 1.0:
            flat_g.mul_(scale)
        out = torch.empty(self.shard_size, dtype=torch.float32, device=self.device)
        dist.reduce_scatter_tensor(out, flat_g, group=self.group)
        self.main_grad_shard.add_(out)

    def copy_master_to_shard(self):
        self.shard.copy_(self.master.data)


class FullyShardedDataParallel(torch.nn.Module):
    def __init__(self, config, ddp_config, module: torch.nn.Module, fsdp_unit_modules: Optional[Sequence[type]] = None, group=None,
                 disable_bucketing: bool = False, data_parallel_sharding_strategy: Optional[str] = None, outer_dp_group=None, prefetch: bool = True, **_):
        """``data_parallel_sharding_strategy`` (reference ``--data-parallel-sharding-strategy``): ``optim_grads_params`` = ZeRO-3 (parameters released
        outside their unit's compute), ``optim_grads`` = ZeRO-2 and ``optim`` = ZeRO-1 (parameters stay resident, only the all-gather after the
        optimizer step remains; gradients still land reduce-scattered on the shard, which is what the sharded optimizer consumes).
        ``outer_dp_group`` enables HSDP: shard inside ``group`` (one NVLink island), replicate across ``outer_dp_group`` — shard gradients are
        all-reduced over the outer group once per step."""
        super().__init__()
        self.config, self.ddp_config, self.module = config, ddp_config, module
        self.group = group if group is not None else ps.get_data_parallel_group(with_context_parallel=True)
        self.world = dist.get_world_size(self.group)
        self.strategy = data_parallel_sharding_strategy or getattr(ddp_config, "data_parallel_sharding_strategy", None) or "optim_grads_params"
        assert self.strategy in ("optim", "optim_grads", "optim_grads_params", "no_shard"), self.strategy
        self.release_params = self.strategy == "optim_grads_params"
        self.outer_group = outer_dp_group
        self.outer_world = dist.get_world_size(outer_dp_group) if outer_dp_group is not None else 1
        self.prefetch = prefetch and self.release_params
        if fsdp_unit_modules is None:
            from ...transformer.transformer_layer import TransformerLayer

            fsdp_unit_modules = (TransformerLayer,)
        self.units: List[_Unit] = []
        self._unit_of_module: Dict[torch.nn.Module, _Unit] = {}
        claimed = set()
        ordered: List[tuple] = []
        for name, sub in module.named_modules():
            if isinstance(sub, tuple(fsdp_unit_modules)):
                ps_ = [p for p in sub.parameters() if p.requires_grad and id(p) not in claimed]
                if ps_:
                    claimed.update(id(p) for p in ps_)
                    ordered.append((name, sub, ps_))
        root = [p for p in module.parameters() if p.requires_grad and id(p) not in claimed]
        if root:
            ordered.insert(0, ("<root>", module, root))
        for name, sub, ps_ in ordered:
            # one unit per dtype keeps the flat buffer homogeneous
            by_dtype: Dict[torch.dtype, List[torch.nn.Parameter]] = {}
            for p in ps_:
                by_dtype.setdefault(p.dtype, []).append(p)
            for dt, plist in by_dtype.items():
                u = _Unit(f"{name}[{dt}]", plist, self.group)
                self.units.append(u)
                self._unit_of_module.setdefault(sub, u)
                self._hook_unit(sub, u, is_root=name == "<root>")
        self._scale = 1.0 / (self.world * self.outer_world)
        self._sync = True
        self._layer_units = [u for u in self.units if not getattr(u, "is_root", False)]
        for i, u in enumerate(self._layer_units):
            u.index = i
        if self.release_params:
            for u in self._layer_units:
                u.release()

    # ---- hooks ----------------------------------------------------------------------------------------------------
    def _hook_unit(self, sub: torch.nn.Module, u: _Unit, is_root: bool):
        def neighbour(step):
            i = getattr(u, "index", None)
            if i is None or not self.prefetch:
                return None
            j = i + step
            return self._layer_units[j] if 0 <= j < len(self._layer_units) else None

        def pre_fwd(mod, args):
            u.gather()
            u.wait()
            nxt = neighbour(+1)
            if nxt is not None:
                nxt.gather(async_op=True)      # overlaps with this unit's compute

        def post_fwd(mod, args, out):
            if not is_root and self.release_params:
                u.release()
            return out

        def pre_bwd(mod, gout):
            u.gather()
            u.wait()
            u.pending = sum(1 for p in u.params if p.requires_grad)
            prv = neighbour(-1)
            if prv is not None:
                prv.gather(async_op=True)

        sub.register_forward_pre_hook(pre_fwd)
        if not is_root:
            sub.register_forward_hook(post_fwd)
            sub.register_full_backward_pre_hook(pre_bwd)

        def on_grad(p):
            u.pending -= 1
            if u.pending == 0 and not is_root:
                u.reduce_scatter_grads(self._scale)
                if self.release_params:
                    u.release()

        for p in u.params:
            p.register_post_accumulate_grad_hook(on_grad)
        u.is_root = is_root

    # ---- DDP-like API used by the training loop ----------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        for u in self.units:
            if getattr(u, "is_root", False):
                u.gather()
                u.wait()
                u.pending = len(u.params)
        return self.module(*args, **kwargs)

    def finish_grad_sync(self, force_all_reduce: bool = False):
        """Root-unit gradients (embedding, final norm, head) are reduced here; layer units were reduced in backward."""
        for u in self.units:
            if getattr(u, "is_root", False):
                u.reduce_scatter_grads(self._scale)
            if self.outer_group is not None and self.outer_world > 1:
                dist.all_reduce(u.main_grad_shard, group=self.outer_group)     # HSDP: replicas of this shard live on the other islands
            u.master.grad = u.main_grad_shard

    start_grad_sync = finish_grad_sync

    def zero_grad_buffer(self):
        for u in self.units:
            u.main_grad_shard.zero_()
            u.master.grad = None

    @contextmanager
    def no_sync(self):
        yield  # ZeRO-3 reduces every micro-batch into the fp32 shard accumulator; nothing to defer

    def optimizer_parameters(self) -> List[torch.nn.Parameter]:
        """fp32 master shards — what the optimizer steps on."""
        return [u.master for u in self.units]

    def post_optimizer_step(self):
        """Copy updated master shards back into the model-dtype shards; root params are re-gathered eagerly."""
        for u in self.units:
            u.copy_master_to_shard()
            if u.resident:
                u.resident = False  # contents stale
                u.gather()
                u.wait()

    def gather_full_state_dict(self) -> Dict[str, torch.Tensor]:
        """Materialise full (unsharded) parameters on every rank — checkpoint export / tests."""
        for u in self.units:
            u.gather()
            u.wait()
        sd = {k: v.detach().clone() for k, v in self.module.state_dict().items()}
        for u in self.units:
            if not getattr(u, "is_root", False) and self.release_params:
                u.release()
        return sd

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def scale_gradients(self, factor: float) -> None:
        """Per-token loss normalisation (``finalize_model_grads`` divides by the global token count)."""
        for u in self.units:
            u.main_grad_shard.mul_(factor)

    def broadcast_params(self) -> None:
        """Make every data-parallel rank start from rank 0's parameters: each unit's shard is a slice of the SAME flat vector only if the ranks initialised
        identically — here rank 0's full flat buffer is broadcast and re-sliced."""
        for u in self.units:
            was = u.resident
            u.gather()
            u.wait()
            dist.broadcast(u.flat, src=dist.get_global_rank(self.group, 0), group=self.group)
            lo = u.rank * u.shard_size
            u.shard.copy_(u.flat[lo:lo + u.shard_size])
            u.master.data.copy_(u.shard)
            if not was and self.release_params and not getattr(u, "is_root", False):
                u.release()

    def load_state_dict(self, sd, strict: bool = True):
        """Accepts either this wrapper's sharded form (``fsdp.unit{i}.master`` → master shards) or a full module state dict."""
        if any(k.startswith("fsdp.unit") or ".fsdp.unit" in k for k in sd):
            for i, u in enumerate(self.units):
                key = next((k for k in sd if k.endswith(f"fsdp.unit{i}.master")), None)
                if key is None:
                    if strict:
                        raise KeyError(f"fsdp.unit{i}.master missing from the checkpoint")
                    continue
                u.master.data.copy_(sd[key])
            self.post_optimizer_step()
            return
        for u in self.units:
            u.gather()
            u.wait()
        out = self.module.load_state_dict(sd, strict=strict)
        for u in self.units:
            lo = u.rank * u.shard_size
            u.shard.copy_(u.flat[lo:lo + u.shard_size])
            u.master.data.copy_(u.shard)
            if self.release_params and not getattr(u, "is_root", False):
                u.release()
        return out

    def sharded_state_dict(self, prefix: str = "", *a, **k):
        from ...dist_checkpointing.mapping import ShardedTensor

        out = {}
        for i, u in enumerate(self.units):
            key = f"{prefix}fsdp.unit{i}.master"
            out[key] = ShardedTensor.from_rank_offsets(key, u.master.data, (0, u.rank, u.world), replica_id=0)
        return out
