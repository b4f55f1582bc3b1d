"""ZeRO-1 distributed optimizer (reference ``optimizer/distrib_optimizer.py:113``).

Each data-parallel rank owns the ``[r·n/dp, (r+1)·n/dp)`` range of every bucket of the
DDP grad/param buffers: after the bucket's reduce-scatter that range of ``grad_data``
holds the averaged gradient, the fused Adam kernel updates the fp32 master shard and
writes the bf16 result straight into the same range of ``param_data``, and the bucket's
all-gather (NVLink multicast store on B200: ``NVLinkBackend.all_gather_inplace_``)
publishes it to the other ranks.

Checkpoint formats (``metadata['distrib_optim_sharding_type']``):
  * ``dp_reshardable``      flat per-bucket shards (fast; only DP may change)
  * ``fully_reshardable``   model-shaped tensors sharded like the weight (TP/PP/DP may change)
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import parallel_state as ps
from ..dist_checkpointing.mapping import ShardedTensor
from ..utils import get_pg_rank, get_pg_size
from .grad_scaler import MegatronGradScaler
from .optimizer import MixedPrecisionOptimizer, Slot
from .optimizer_config import OptimizerConfig


class Range:
    """Half-open integer interval."""

    def __init__(self, start: int, end: int):
        self.start, self.end = start, end
        self.size = end - start

    def normalize(self, start: int = 0) -> "Range":
        return Range(start, start + self.size)

    def __len__(self):
        return self.size

    def __repr__(self):
        return f"[{self.start},{self.end})"


class DistributedOptimizer(MixedPrecisionOptimizer):
    def __init__(self, param_groups: List[dict], config: OptimizerConfig, grad_scaler: Optional[MegatronGradScaler],
                 init_state_fn: Optional[Callable], model_chunks: List, per_model_buffers: Dict[int, List], data_parallel_group,
                 data_parallel_group_gloo=None, data_parallel_group_idx: int = 0, distributed_optimizer_instance_id: int = 0):
        super().__init__(param_groups, config, grad_scaler, init_state_fn or (lambda x: None))
        self.model_chunks = model_chunks
        self.ddp_config = model_chunks[0].ddp_config if model_chunks else None
        self.per_model_buffers = per_model_buffers
        self.data_parallel_group = data_parallel_group
        self.data_parallel_group_gloo = data_parallel_group_gloo
        self.data_parallel_group_idx = data_parallel_group_idx
        self.dp_size = get_pg_size(data_parallel_group)
        self.dp_rank = get_pg_rank(data_parallel_group)
        self.buffers = [b for idx in sorted(per_model_buffers) for b in per_model_buffers[idx]]
        param_to_group = {p: gi for gi, g in enumerate(param_groups) for p in g["params"]}

        # ---- range maps → slots ----------------------------------------------------------
        self.param_ranges: Dict[torch.nn.Parameter, Dict[str, Range]] = {}
        self.slot_meta: List[Tuple[int, int, int, Range]] = []  # (buffer_idx, bucket_idx, slot_idx, range in buffer)
        for bi, buf in enumerate(self.buffers):
            assert buf.param_data is not None, "DistributedOptimizer needs DDP built with use_distributed_optimizer=True"
            for bucket in buf.buckets:
                n = bucket.grad_data.numel()
                assert n % self.dp_size == 0, "bucket must be padded to a multiple of the data-parallel size"
                shard = n // self.dp_size
                w0 = bucket.offset + self.dp_rank * shard
                w1 = w0 + shard
                for p in bucket.params_list:
                    s, e, _ = buf.param_index_map[p]
                    lo, hi = max(s, w0), min(e, w1)
                    if hi <= lo or p not in param_to_group:
                        continue
                    self.param_ranges[p] = {"buffer": Range(lo, hi), "param": Range(lo - s, hi - s), "bucket": Range(lo - bucket.offset, hi - bucket.offset)}
                    lowp = buf.param_data[lo:hi] if buf.param_dtype != torch.float32 else None
                    master = buf.param_data[lo:hi].detach().clone().float() if lowp is not None else buf.param_data[lo:hi]
                    slot = Slot(p, (lambda _b=buf, _lo=lo, _hi=hi: _b.grad_data[_lo:_hi]), master, lowp, param_to_group[p], name=buf.param_to_name.get(p))
                    self.slot_meta.append((bi, bucket.bucket_id, len(self.slots), Range(lo, hi)))
                    self.slots.append(slot)
        self.is_stub_optimizer = len(self.slots) == 0 and not any(g["params"] for g in param_groups)

    # the grad norm of a sharded optimizer is reduced over model-parallel × this DP group
    def get_grad_stats_parallel_group(self):
        if self.grad_stats_parallel_group is not None:
            return self.grad_stats_parallel_group
        if not ps.is_initialized():
            return None
        # every rank holds a disjoint (model-parallel shard × dp shard) piece ⇒ reduce over the world
        return dist.group.WORLD

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None
        # param all-gather for the *next* forward is dispatched here when overlapped
        if self.ddp_config is not None and self.ddp_config.overlap_param_gather and getattr(self, "_params_dirty", False):
            for c in self.model_chunks:
                c.start_param_sync()
            self._params_dirty = False

    def _post_step(self):
        """Publish the updated bf16 shards to all data-parallel ranks."""
        if self.dp_size == 1:
            return
        if self.ddp_config.overlap_param_gather:
            self._params_dirty = True
            return
        timers = self.config.timers
        if timers is not None:
            timers("params-all-gather", log_level=1).start(barrier=self.config.barrier_with_L1_time)
        for c in self.model_chunks:
            c.start_param_sync(force_sync=True)
        if timers is not None:
            timers("params-all-gather").stop()

    @torch.no_grad()
    def step(self):
        out = super().step()
        if out[0]:
            self._post_step()
        return out

    def reload_model_params(self, state_dict=None):
        for s in self.slots:
            if s.lowp is not None:
                s.master.copy_(s.lowp)

    # ---- plain state dict (same-layout resume) -------------------------------------------
    def state_dict(self):
        return {
            "optimizer": {"param_groups": self._groups_meta()},
            "grad_scaler": self.grad_scaler.state_dict() if self.grad_scaler else None,
            "shards": [{k: v.cpu() for k, v in self._slot_state(s).items()} for s in self.slots],
        }

    def load_state_dict(self, state_dict):
        for gi, meta in enumerate(state_dict["optimizer"]["param_groups"]):
            self.step_count[gi] = meta.get("step", 0)
            for k, v in meta.items():
                if k not in ("params", "step"):
                    self.param_groups[gi][k] = v
        if self.grad_scaler and state_dict.get("grad_scaler"):
            self.grad_scaler.load_state_dict(state_dict["grad_scaler"])
        for s, st in zip(self.slots, state_dict.get("shards", [])):
            self._init_slot_state(s)
            for k, v in st.items():
                (s.master if k == "param" else getattr(s, k)).copy_(v)
            if s.lowp is not None:
                s.lowp.copy_(s.master)

    # ---- sharded state dicts ---------------------------------------------------------------
    def sharded_state_dict(self, model_sharded_state_dict, is_loading: bool = False, metadata: Optional[dict] = None):
        sharding_type = (metadata or {}).get("distrib_optim_sharding_type", "fully_reshardable")
        if sharding_type in ("fully_sharded_model_space", "fully_reshardable"):
            sd = self._sharded_fully_reshardable(model_sharded_state_dict, is_loading)
        elif sharding_type in ("dp_reshardable", "dp_zero_gather_scatter", "fully_sharded_bucket_space"):
            sd = self._sharded_dp_reshardable(is_loading)
        else:
            raise NotImplementedError(f"unknown optimizer sharding type {sharding_type}")
        sd["param_state_sharding_type"] = sharding_type
        return sd

    def _common(self):
        return {"optimizer": {"param_groups": self._groups_meta()}, "grad_scaler": self.grad_scaler.state_dict() if self.grad_scaler else None}

    def _sharded_dp_reshardable(self, is_loading: bool):
        """One flat ShardedTensor per (buffer, bucket, state): global = padded bucket, fragmented dp ways."""
        if is_loading:
            for s in self.slots:
                self._init_slot_state(s)
        out = self._common()
        state = {}
        names = ["param", "exp_avg", "exp_avg_sq"] if self.config.optimizer == "adam" else ["param", "momentum"]
        by_bucket: Dict[Tuple[int, int], List[Tuple[Slot, Range]]] = {}
        for bi, bid, si, rng in self.slot_meta:
            by_bucket.setdefault((bi, bid), []).append((self.slots[si], rng))
        mp_rank = dist.get_rank(ps.get_model_parallel_group()) if ps.is_initialized() else 0
        for (bi, bid), lst in by_bucket.items():
            buf = self.buffers[bi]
            bucket = buf.buckets[bid]
            shard = bucket.grad_data.numel() // self.dp_size
            w0 = bucket.offset + self.dp_rank * shard
            for nm in names:
                flat = torch.zeros(shard, dtype=torch.float32, device=bucket.grad_data.device)
                for s, rng in lst:
                    src = s.master if nm == "param" else getattr(s, nm)
                    if src is not None:
                        flat[rng.start - w0 : rng.end - w0] = src
                key = f"optimizer.distributed.dp_group_idx_{self.data_parallel_group_idx}.mp_{mp_rank}.gbuf_idx_{bi}.dtype_{buf.param_dtype}_{buf.grad_dtype}.bucket_idx_{bid}.{nm}"
                state[(bi, bid, nm)] = ShardedTensor.from_rank_offsets(key, flat, (0, self.dp_rank, self.dp_size), replica_id=0)
        out["param_state"] = {f"{a}.{b}.{c}": v for (a, b, c), v in state.items()}
        self._dp_reshardable_layout = by_bucket
        return out

    def _gather_full(self, src_of_slot, templates_only: bool = False) -> Dict[torch.nn.Parameter, torch.Tensor]:
        """Model-shaped fp32 tensors of one optimizer state, staged on the HOST.

        One all-gather per bucket (its dp shards are contiguous slices of the bucket), never a full-model buffer: the
        device high-water mark is one bucket, and nothing stays resident after the checkpoint call (the reference gathers
        per bucket to dp rank 0 as well, ``distrib_optimizer.py:1300-1420``).  ``templates_only`` (loading): no
        communication, just empty destinations."""
        full: Dict[torch.nn.Parameter, torch.Tensor] = {}
        pin = torch.cuda.is_available()
        if templates_only:
            for buf in self.buffers:
                for p in buf.param_index_map:
                    full[p] = torch.empty(p.shape, dtype=torch.float32)
            return full
        pieces: Dict[Tuple[int, int], List[Tuple[Slot, Range]]] = {}
        for bi, bid, si, rng in self.slot_meta:
            pieces.setdefault((bi, bid), []).append((self.slots[si], rng))
        for bi, buf in enumerate(self.buffers):
            params_of_bucket: Dict[int, list] = {}
            for p, (s, e, bid) in buf.param_index_map.items():
                params_of_bucket.setdefault(bid, []).append((p, s, e))
            for bid, bucket in enumerate(buf.buckets):
                n = bucket.grad_data.numel()
                shard = n // self.dp_size
                w0 = bucket.offset + self.dp_rank * shard
                mine = torch.zeros(shard, dtype=torch.float32, device=bucket.grad_data.device)
                for slot, rng in pieces.get((bi, bid), []):
                    src = src_of_slot(slot)
                    if src is not None:
                        mine[rng.start - w0 : rng.end - w0] = src
                if self.dp_size > 1:
                    flat = torch.empty(n, dtype=torch.float32, device=mine.device)
                    dist.all_gather_into_tensor(flat, mine, group=self.data_parallel_group)
                else:
                    flat = mine
                for p, s, e in params_of_bucket.get(bid, []):
                    host = torch.empty(p.shape, dtype=torch.float32, pin_memory=pin)
                    host.copy_(flat[s - bucket.offset : e - bucket.offset].view(p.shape))
                    full[p] = host
                del flat, mine
        return full

    def _sharded_fully_reshardable(self, model_sharded_state_dict, is_loading: bool):
        from ..dist_checkpointing.optimizer import get_param_id_to_sharded_param_map, make_sharded_optimizer_tensor

        for s in self.slots:
            self._init_slot_state(s)
        names = ["param", "exp_avg", "exp_avg_sq"] if self.config.optimizer == "adam" else ["param", "momentum"]
        params = [p for buf in self.buffers for p in buf.param_index_map]
        id_map = get_param_id_to_sharded_param_map(model_sharded_state_dict, params)
        out = self._common()
        st = {}
        for nm in names:
            full = self._gather_full(lambda s, _nm=nm: s.master if _nm == "param" else getattr(s, _nm), templates_only=is_loading)
            for i, p in enumerate(params):
                key_nm = "fp32_param" if nm == "param" else nm
                st.setdefault(i, {})[key_nm] = make_sharded_optimizer_tensor(id_map[i], full[p], prefix=f"optimizer.state.{key_nm}")
        out["param_state"] = st
        self._full_params_order = params
        return out

    def load_sharded_state_dict(self, sd):
        for gi, meta in enumerate(sd["optimizer"]["param_groups"]):
            self.step_count[gi] = meta.get("step", 0)
        if self.grad_scaler and sd.get("grad_scaler"):
            self.grad_scaler.load_state_dict(sd["grad_scaler"])
        kind = sd.get("param_state_sharding_type", "fully_reshardable")
        for s in self.slots:
            self._init_slot_state(s)
        if kind in ("fully_sharded_model_space", "fully_reshardable"):
            params = self._full_params_order
            for i, p in enumerate(params):
                rng = self.param_ranges.get(p)
                if rng is None:
                    continue
                slot = next(s for s in self.slots if s.param is p)
                for key_nm, t in sd["param_state"][i].items():
                    dst = slot.master if key_nm == "fp32_param" else getattr(slot, key_nm)
                    dst.copy_(t.reshape(-1)[rng["param"].start : rng["param"].end])
        else:
            for (bi, bid), lst in self._dp_reshardable_layout.items():
                bucket = self.buffers[bi].buckets[bid]
                shard = bucket.grad_data.numel() // self.dp_size
                w0 = bucket.offset + self.dp_rank * shard
                for s, rng in lst:
                    for nm in ("param", "exp_avg", "exp_avg_sq", "momentum"):
                        t = sd["param_state"].get(f"{bi}.{bid}.{nm}")
                        if t is None:
                            continue
                        dst = s.master if nm == "param" else getattr(s, nm)
                        if dst is not None:
                            dst.copy_(t[rng.start - w0 : rng.end - w0])
        for s in self.slots:
            if s.lowp is not None:
                s.lowp.copy_(s.master)
        for c in self.model_chunks:
            if self.dp_size > 1:
                c.start_param_sync(force_sync=True)
