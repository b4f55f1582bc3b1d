"""Glue between the sharded-parameter wrapper and the training loop (reference ``distributed/fsdp/mcore_fsdp_adapter.py:66-760``): the pieces the loop expects
from a DDP-wrapped model + Megatron optimizer pair, implemented on FSDP units.

* ``FSDPOptimizer`` — steps the fp32 MASTER SHARDS of every unit (that is all the optimizer state a rank holds: ZeRO-1/2/3 differ only in what else stays
  resident), computes the global gradient norm from the shard gradients with one all-reduce, clips, then lets every unit copy master → model-dtype shard (the
  all-gather of the next forward distributes the update).  Exposes ``param_groups`` (LR scheduler), ``state_dict`` / ``sharded_state_dict`` (per-unit master +
  Adam moments as 1-D tensors sharded over the data-parallel group) and the no-op loss-scaling surface of bf16 training.
* ``fsdp_model_state_dict`` helpers — the wrapper's own ``sharded_state_dict`` / ``load_state_dict`` keep a checkpoint loadable at the same world size; the
  model-space format in ``transformer/fsdp_dtensor_checkpoint.py`` reshards across world sizes."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .fully_sharded_data_parallel import FullyShardedDataParallel


class FSDPOptimizer:
    is_stub_optimizer = False

    def __init__(self, models: List[FullyShardedDataParallel], config, model_parallel_group=None):
        self.models, self.config, self.mp_group = models, config, model_parallel_group
        params = [p for m in models for p in m.optimizer_parameters()]
        kind = getattr(config, "optimizer", "adam")
        lr, wd = getattr(config, "lr", 1e-3) or 1e-3, getattr(config, "weight_decay", 0.0) or 0.0
        if kind == "sgd":
            self.optimizer = torch.optim.SGD(params, lr=lr, momentum=getattr(config, "sgd_momentum", 0.0), weight_decay=wd)
        else:
            self.optimizer = torch.optim.AdamW(params, lr=lr, betas=(getattr(config, "adam_beta1", 0.9), getattr(config, "adam_beta2", 0.999)),
                                               eps=getattr(config, "adam_eps", 1e-8), weight_decay=wd)
        for g in self.optimizer.param_groups:      # the scheduler multiplies these
            g.setdefault("lr_mult", 1.0)
            g.setdefault("wd_mult", 1.0)
            g.setdefault("is_expert_parallel", False)
            g.setdefault("max_lr", lr)
            g.setdefault("min_lr", getattr(config, "min_lr", 0.0) or 0.0)
        self.clip_grad = getattr(config, "clip_grad", 0.0) or 0.0
        self.step_count = 0

    # ---- the MegatronOptimizer surface used by train_step / the schedules -------------------------------------------------------------
    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def zero_grad(self, set_to_none: bool = True):
        for m in self.models:
            m.zero_grad_buffer()

    def get_loss_scale(self):
        return torch.ones(1)

    def scale_loss(self, loss):
        return loss

    def reload_model_params(self, state_dict=None):
        for m in self.models:
            for u in m.units:
                u.master.data.copy_(u.shard)

    def get_grad_norm(self) -> float:
        total = torch.zeros(1, dtype=torch.float32, device=self.models[0].units[0].master.device)
        for m in self.models:
            for u in m.units:
                if u.master.grad is not None:
                    total += u.master.grad.float().pow(2).sum()
        dist.all_reduce(total, group=self.models[0].group)              # shards are disjoint across the data-parallel group
        if self.mp_group is not None and dist.get_world_size(self.mp_group) > 1:
            dist.all_reduce(total, group=self.mp_group)
        return float(total.sqrt())

    @torch.no_grad()
    def step(self):
        grad_norm = self.get_grad_norm()
        if not (grad_norm == grad_norm and grad_norm != float("inf")):      # NaN / inf: skip the update on every rank (the norm is global)
            return False, grad_norm, None
        if self.clip_grad > 0 and grad_norm > self.clip_grad:
            coef = self.clip_grad / (grad_norm + 1e-6)
            for m in self.models:
                for u in m.units:
                    if u.master.grad is not None:
                        u.master.grad.mul_(coef)
        self.optimizer.step()
        for m in self.models:
            m.post_optimizer_step()
        self.step_count += 1
        return True, grad_norm, None

    # ---- state ---------------------------------------------------------------------------------------------------------------------------
    def state_dict(self):
        return {"optimizer": self.optimizer.state_dict(), "step_count": self.step_count}

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd["optimizer"])
        self.step_count = sd.get("step_count", 0)

    def sharded_state_dict(self, model_sharded_state_dict=None, is_loading: bool = False, **_) -> Dict:
        from ...dist_checkpointing.mapping import ShardedObject, ShardedTensor

        out: Dict = {}
        k = 0
        for mi, m in enumerate(self.models):
            for ui, u in enumerate(m.units):
                st = self.optimizer.state.get(u.master, {})
                for name in ("exp_avg", "exp_avg_sq", "momentum_buffer"):
                    if name in st or (is_loading and name != "momentum_buffer" and isinstance(self.optimizer, torch.optim.AdamW)):
                        t = st.get(name)
                        if t is None:
                            t = torch.zeros_like(u.master.data)
                            self.optimizer.state.setdefault(u.master, {})[name] = t
                        key = f"optimizer.fsdp.m{mi}.unit{ui}.{name}"
                        out[key] = ShardedTensor.from_rank_offsets(key, t, (0, u.rank, u.world), replica_id=0)
                k += 1
        steps = [float(self.optimizer.state[p]["step"]) if "step" in self.optimizer.state.get(p, {}) else 0.0 for m in self.models for p in m.optimizer_parameters()]
        rank = dist.get_rank() if dist.is_initialized() else 0
        out["optimizer.fsdp.meta"] = ShardedObject("optimizer.fsdp.meta", {"steps": steps, "step_count": self.step_count, "param_groups": [
            {kk: vv for kk, vv in g.items() if kk != "params"} for g in self.optimizer.param_groups]}, (1,), (0,), replica_id=rank)
        return out

    def load_sharded_state_dict(self, loaded: Dict) -> None:
        meta = loaded.get("optimizer.fsdp.meta") or {}
        params = [p for m in self.models for p in m.optimizer_parameters()]
        for p, s in zip(params, meta.get("steps", [])):
            if p in self.optimizer.state:
                self.optimizer.state[p]["step"] = torch.tensor(float(s))
        for g, saved in zip(self.optimizer.param_groups, meta.get("param_groups", [])):
            g.update({k: v for k, v in saved.items() if k in ("lr", "weight_decay", "betas", "eps")})
        self.step_count = meta.get("step_count", self.step_count)
        # the moment tensors were filled in place by dist_checkpointing.load (they are the optimizer's own state tensors)


def wrap_model_with_fsdp(chunks: List[torch.nn.Module], config, ddp_config, strategy: Optional[str] = None, group=None, outer_dp_group=None) -> List[FullyShardedDataParallel]:
    return [FullyShardedDataParallel(config, ddp_config, c, data_parallel_sharding_strategy=strategy, group=group, outer_dp_group=outer_dp_group) for c in chunks]
