"""Layer-wise distributed optimizer (reference ``optimizer/layer_wise_optimizer.py`` — ``LayerWiseDistributedOptimizer``).

ZeRO-1 shards FLAT ranges of the parameter buffer, which only works for element-wise update rules.  Optimizers that need whole matrices (Muon's
Newton-Schulz, SOAP / Shampoo preconditioners) are distributed by OWNERSHIP instead: every parameter is assigned to exactly one data-parallel
rank (longest-processing-time packing on the element count, so optimizer memory and time are balanced), gradients are all-reduced as in plain DDP,
each rank updates the parameters it owns — its optimizer state exists nowhere else — and the updated parameters are broadcast from their owners,
coalesced into ONE flat buffer per owner (``world`` broadcasts per step, independent of the number of parameters).

The wrapped optimizer is one of this package's mixed-precision / fp32 optimizers built over the owned parameters only, so fp32 masters, loss
scaling, clipping (with the GLOBAL gradient norm), Muon / Lion / Adam update rules and checkpointing are inherited."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .optimizer import Float16OptimizerWithFloat16Params, FP32Optimizer
from .optimizer_config import OptimizerConfig


def partition_params_by_owner(named_params: List[tuple], world: int) -> Dict[str, int]:
    """name → owning rank.  Deterministic on every rank: sort by (-numel, name), give each to the currently lightest rank."""
    load = [0] * world
    owner = {}
    for name, p in sorted(named_params, key=lambda kv: (-kv[1].numel(), kv[0])):
        r = min(range(world), key=lambda i: (load[i], i))
        owner[name] = r
        load[r] += p.numel()
    return owner


class LayerWiseDistributedOptimizer:
    def __init__(self, config: OptimizerConfig, model_chunks: List[torch.nn.Module], param_groups: List[dict], data_parallel_group=None, grad_scaler=None,
                 model_parallel_group=None):
        self.config, self.model_chunks, self.dp_group = config, model_chunks, data_parallel_group
        self.world = dist.get_world_size(data_parallel_group) if data_parallel_group is not None else 1
        self.rank = dist.get_rank(data_parallel_group) if data_parallel_group is not None else 0
        named = []
        for ci, m in enumerate(model_chunks):
            named += [(f"{ci}.{n}", p) for n, p in m.named_parameters() if p.requires_grad]
        self.name_of = {id(p): n for n, p in named}
        self.owner = partition_params_by_owner(named, self.world)
        self.params_by_owner: List[List[torch.nn.Parameter]] = [[] for _ in range(self.world)]
        for n, p in sorted(named, key=lambda kv: kv[0]):
            self.params_by_owner[self.owner[n]].append(p)
        owned_groups = []
        for g in param_groups:
            mine = [p for p in g["params"] if p.requires_grad and self.owner[self.name_of[id(p)]] == self.rank]
            owned_groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": mine})
        for g in owned_groups:
            g.setdefault("betas", (config.adam_beta1, config.adam_beta2))
            g.setdefault("eps", config.adam_eps)
        lowp = config.fp16 or config.bf16
        self.inner = Float16OptimizerWithFloat16Params(owned_groups, config, grad_scaler) if lowp else FP32Optimizer(owned_groups, config)
        self.inner.model_chunks = model_chunks
        if model_parallel_group is not None:
            self.inner.grad_stats_parallel_group = model_parallel_group
        self.param_groups = self.inner.param_groups
        self.is_stub_optimizer = False

    # ---- delegation ----
    def zero_grad(self, set_to_none: bool = True):
        for m in self.model_chunks:
            for p in m.parameters():
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def get_loss_scale(self):
        return self.inner.get_loss_scale()

    def scale_loss(self, loss):
        return self.inner.scale_loss(loss)

    def state_dict(self):
        return {"owner": dict(self.owner), "inner": self.inner.state_dict()}

    def load_state_dict(self, sd):
        assert sd["owner"] == self.owner, "layer-wise ownership changed (different DP size): reshard through sharded_state_dict instead"
        self.inner.load_state_dict(sd["inner"])

    def sharded_state_dict(self, model_sharded_state_dict, is_loading: bool = False, metadata: Optional[dict] = None):
        return self.inner.sharded_state_dict(model_sharded_state_dict, is_loading, metadata)

    def load_sharded_state_dict(self, sd):
        return self.inner.load_sharded_state_dict(sd)

    # ---- the step ----
    def _global_grad_norm(self) -> torch.Tensor:
        local = self.inner.get_grad_norm().float() if self.inner.slots else torch.zeros((), device=self._device())
        sq = (local * local).reshape(1).to(self._device())
        if self.world > 1:
            dist.all_reduce(sq, group=self.dp_group)
        return sq.sqrt().reshape(())

    def _device(self):
        p = next(self.model_chunks[0].parameters())
        return p.device

    def get_grad_norm(self):
        return self._global_grad_norm()

    @torch.no_grad()
    def step(self):
        gn = self._global_grad_norm()                              # clip (and detect overflow) with the norm over ALL parameters, not just the owned ones
        self.inner.get_grad_norm = lambda: gn
        try:
            ok, grad_norm, num_zeros = self.inner.step()
        finally:
            del self.inner.get_grad_norm
        if self.world > 1:
            flag = torch.tensor([1.0 if ok else 0.0], device=self._device())
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.dp_group)
            ok = bool(flag.item() > 0)
            self.broadcast_params()
        return ok, grad_norm, num_zeros

    @torch.no_grad()
    def broadcast_params(self) -> None:
        """One coalesced broadcast per owner."""
        ranks = dist.get_process_group_ranks(self.dp_group)
        for r, plist in enumerate(self.params_by_owner):
            if not plist:
                continue
            by_dtype: Dict[torch.dtype, List[torch.nn.Parameter]] = {}
            for p in plist:
                by_dtype.setdefault(p.dtype, []).append(p)
            for dt, ps_ in by_dtype.items():
                flat = torch.cat([p.data.reshape(-1) for p in ps_]) if r == self.rank else torch.empty(sum(p.numel() for p in ps_), dtype=dt, device=ps_[0].device)
                dist.broadcast(flat, src=ranks[r], group=self.dp_group)
                if r != self.rank:
                    off = 0
                    for p in ps_:
                        p.data.copy_(flat[off : off + p.numel()].view_as(p))
                        off += p.numel()

    def optimizer_memory_elements(self) -> List[int]:
        """Elements owned per rank (balance diagnostic)."""
        return [sum(p.numel() for p in pl) for pl in self.params_by_owner]
