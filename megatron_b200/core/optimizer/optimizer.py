"""Mixed-precision optimizers (reference ``optimizer/optimizer.py``: ``MegatronOptimizer`` :187,
``MixedPrecisionOptimizer`` :654, ``Float16OptimizerWithFloat16Params`` :964, ``FP32Optimizer``
:1264, ``ChainedOptimizer`` :1451).

Design: every optimizer owns a list of *slots* — (gradient view, fp32 master, moments,
low-precision destination view).  ``step()`` is

    norm kernel → (all-reduce of one scalar) → clip coefficient on device
    → ONE fused multi-tensor kernel per param group: unscale·clip, Adam(W) on the fp32
      master, write the bf16 model copy

so gradients and states are each read once and nothing syncs with the host unless fp16
loss scaling needs the found-inf flag.  The reference does copy-to-main-grad, unscale,
norm, clip (scale pass), Adam, copy-to-model as separate passes.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from ... import ops
from .. import parallel_state as ps
from ..tensor_parallel import param_is_not_tensor_parallel_duplicate
from ..transformer.module import param_is_not_shared
from ..utils import get_pg_size
from .clip_grads import clip_coefficient, get_grad_norm_fp32
from .grad_scaler import MegatronGradScaler
from .optimizer_config import OptimizerConfig


class Slot:
    """A contiguous run of elements updated together."""

    __slots__ = ("param", "grad_fn", "master", "lowp", "exp_avg", "exp_avg_sq", "momentum", "group", "name", "extra")

    def __init__(self, param, grad_fn, master, lowp, group, name=None):
        self.param, self.grad_fn, self.master, self.lowp, self.group, self.name = param, grad_fn, master, lowp, group, name
        self.exp_avg = self.exp_avg_sq = self.momentum = None
        self.extra = None        # rule-specific matrices (SOAP: Kronecker statistics and their eigenbases)

    @property
    def grad(self):
        return self.grad_fn()


def _zero_grad_group(params, set_to_none=True):
    for p in params:
        if p.grad is not None:
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_().zero_()


class MegatronOptimizer(ABC):
    """Interface shared by all optimizers of this framework."""

    def __init__(self, param_groups: List[dict], config: OptimizerConfig, init_state_fn: Callable = lambda x: None):
        self.param_groups = param_groups
        self.config = config
        self.init_state_fn = init_state_fn
        self.grad_stats_parallel_group = None
        self.tp_group = None
        self.is_stub_optimizer = not any(g["params"] for g in param_groups)
        self.slots: List[Slot] = []
        self.step_count = [0] * len(param_groups)
        self._grad_norm = None

    # compat: callers reach for ``.optimizer.param_groups``
    @property
    def optimizer(self):
        return self

    def get_parameters(self) -> List[torch.nn.Parameter]:
        return [p for g in self.param_groups for p in g["params"]]

    def get_main_grads_for_grad_norm(self) -> List[torch.Tensor]:
        out = []
        for s in self.slots:
            g = s.grad
            if g is None:
                continue
            if param_is_not_shared(s.param) and param_is_not_tensor_parallel_duplicate(s.param, self.tp_group):
                out.append(g)
        return out

    def get_grad_stats_parallel_group(self):
        if self.grad_stats_parallel_group is not None:
            return self.grad_stats_parallel_group
        return ps.get_model_parallel_group(check_initialized=False)

    def get_grad_norm(self) -> torch.Tensor:
        return get_grad_norm_fp32(self.get_main_grads_for_grad_norm(), grad_stats_parallel_group=self.get_grad_stats_parallel_group())

    def clip_grad_norm(self, clip_grad: float) -> torch.Tensor:
        """Returns the norm; the clip itself is folded into the update kernel."""
        norm = self.get_grad_norm()
        self._clip_coeff = clip_coefficient(norm, clip_grad) if clip_grad > 0 else None
        return norm

    def count_zeros(self) -> torch.Tensor:
        total = torch.zeros((), dtype=torch.float32)
        for s in self.slots:
            g = s.grad
            if g is not None and param_is_not_shared(s.param) and param_is_not_tensor_parallel_duplicate(s.param, self.tp_group):
                total = total.to(g.device) + (g.numel() - torch.count_nonzero(g)).float()
        total = total.reshape(1)
        grp = self.get_grad_stats_parallel_group()
        if grp is not None and get_pg_size(grp) > 1:
            dist.all_reduce(total, group=grp)
        return total[0]

    @abstractmethod
    def zero_grad(self, set_to_none: bool = True):
        ...

    @abstractmethod
    def get_loss_scale(self) -> torch.Tensor:
        ...

    def scale_loss(self, loss: torch.Tensor) -> torch.Tensor:
        return self.get_loss_scale().to(loss.device) * loss

    def reload_model_params(self, state_dict=None):
        pass

    @abstractmethod
    def state_dict(self):
        ...

    @abstractmethod
    def load_state_dict(self, state_dict):
        ...

    @abstractmethod
    def step(self):
        ...

    @abstractmethod
    def sharded_state_dict(self, model_sharded_state_dict, is_loading: bool = False, metadata: Optional[dict] = None):
        ...

    # -- the fused update -------------------------------------------------------------
    def _init_slot_state(self, s: Slot):
        if self.config.optimizer == "adam":
            if s.exp_avg is None:
                s.exp_avg = torch.zeros_like(s.master, dtype=self.config.exp_avg_dtype)
                s.exp_avg_sq = torch.zeros_like(s.master, dtype=self.config.exp_avg_sq_dtype)
        elif self.config.optimizer in ("sgd", "lion"):
            if s.momentum is None:
                s.momentum = torch.zeros_like(s.master)
        elif self.config.optimizer == "muon":
            from .muon import is_muon_param

            if is_muon_param(s.param, s.master):
                if s.momentum is None:
                    s.momentum = torch.zeros_like(s.master)
            elif s.exp_avg is None:        # embeddings, norms, biases, heads: AdamW
                s.exp_avg = torch.zeros_like(s.master, dtype=self.config.exp_avg_dtype)
                s.exp_avg_sq = torch.zeros_like(s.master, dtype=self.config.exp_avg_sq_dtype)
        elif self.config.optimizer == "soap":
            from .soap import init_soap_state, is_soap_param

            if s.exp_avg is None:
                s.exp_avg = torch.zeros_like(s.master, dtype=torch.float32)
                s.exp_avg_sq = torch.zeros_like(s.master, dtype=torch.float32)
            if s.extra is None and is_soap_param(s.param, s.master):
                s.extra = init_soap_state(s.master, self.config.soap_max_precond_dim)

    def _apply_update(self, grad_scale: Optional[torch.Tensor]):
        cfg = self.config
        for gi, group in enumerate(self.param_groups):
            slots = [s for s in self.slots if s.group == gi and s.grad is not None and s.master.numel() > 0]
            # every rank counts every group's step (a dp rank may hold no slot of a group; the saved common "step" must agree)
            self.step_count[gi] += 1
            group["step"] = self.step_count[gi]
            if not slots:
                continue
            lr, wd = group["lr"], group.get("weight_decay", 0.0)
            for s in slots:
                self._init_slot_state(s)
            if cfg.optimizer == "adam":
                b1, b2 = group.get("betas", (cfg.adam_beta1, cfg.adam_beta2))
                ops.fused_adam(
                    [s.master for s in slots], [s.grad for s in slots], [s.exp_avg for s in slots], [s.exp_avg_sq for s in slots],
                    [s.lowp for s in slots], lr=lr, beta1=b1, beta2=b2, eps=group.get("eps", cfg.adam_eps), weight_decay=wd,
                    step=self.step_count[gi], adamw=cfg.decoupled_weight_decay, grad_scale=grad_scale,
                )
            elif cfg.optimizer == "sgd":
                mom = group.get("momentum", cfg.sgd_momentum)
                gs = float(grad_scale) if grad_scale is not None else 1.0
                for s in slots:
                    g = s.grad.float() * gs
                    if wd:
                        g = g + wd * s.master
                    s.momentum.mul_(mom).add_(g)
                    s.master.add_(s.momentum, alpha=-lr)
                    if s.lowp is not None:
                        s.lowp.copy_(s.master)
            elif cfg.optimizer == "lion":
                b1, b2 = group.get("betas", (0.9, 0.99))
                gs = float(grad_scale) if grad_scale is not None else 1.0
                for s in slots:
                    g = s.grad.float() * gs
                    upd = torch.sign(s.momentum * b1 + g * (1 - b1))
                    s.master.mul_(1 - lr * wd).add_(upd, alpha=-lr)
                    s.momentum.mul_(b2).add_(g, alpha=1 - b2)
                    if s.lowp is not None:
                        s.lowp.copy_(s.master)
            elif cfg.optimizer == "muon":
                from .muon import is_muon_param, muon_step

                gs = float(grad_scale) if grad_scale is not None else 1.0
                mu = [s for s in slots if is_muon_param(s.param, s.master)]
                rest = [s for s in slots if not is_muon_param(s.param, s.master)]
                for s in mu:
                    muon_step(s.master, s.grad.float() * gs, s.momentum, lr=lr, weight_decay=wd, config=cfg, param=s.param)
                    if s.lowp is not None:
                        s.lowp.copy_(s.master)
                if rest:
                    b1, b2 = group.get("betas", (cfg.adam_beta1, cfg.adam_beta2))
                    ops.fused_adam([s.master for s in rest], [s.grad for s in rest], [s.exp_avg for s in rest], [s.exp_avg_sq for s in rest], [s.lowp for s in rest],
                                   lr=lr, beta1=b1, beta2=b2, eps=group.get("eps", cfg.adam_eps), weight_decay=wd, step=self.step_count[gi], adamw=True,
                                   grad_scale=grad_scale)
            elif cfg.optimizer == "soap":
                from .soap import soap_step

                gs = float(grad_scale) if grad_scale is not None else 1.0
                b1, b2 = group.get("betas", (cfg.adam_beta1, cfg.adam_beta2))
                so = [s for s in slots if s.extra is not None]
                rest = [s for s in slots if s.extra is None]
                for s in so:
                    soap_step(s.master, s.grad.float() * gs, s.exp_avg, s.exp_avg_sq, s.extra, lr=lr, beta1=b1, beta2=b2, eps=group.get("eps", cfg.adam_eps),
                              weight_decay=wd, step=self.step_count[gi], shampoo_beta=cfg.soap_shampoo_beta,
                              precondition_frequency=cfg.soap_precondition_frequency, precondition_warmup=cfg.soap_precondition_warmup)
                    if s.lowp is not None:
                        s.lowp.copy_(s.master)
                if rest:
                    ops.fused_adam([s.master for s in rest], [s.grad for s in rest], [s.exp_avg for s in rest], [s.exp_avg_sq for s in rest], [s.lowp for s in rest],
                                   lr=lr, beta1=b1, beta2=b2, eps=group.get("eps", cfg.adam_eps), weight_decay=wd, step=self.step_count[gi], adamw=True,
                                   grad_scale=grad_scale)
            else:
                raise NotImplementedError(cfg.optimizer)

    # -- plain (non-sharded) state dict helpers -----------------------------------------
    def _slot_state(self, s: Slot) -> Dict[str, torch.Tensor]:
        d = {"param": s.master}
        for k in ("exp_avg", "exp_avg_sq", "momentum"):
            v = getattr(s, k)
            if v is not None:
                d[k] = v
        if s.extra:
            for k, v in s.extra.items():
                d[f"extra.{k}"] = v if torch.is_tensor(v) else torch.tensor(v)
        return d

    def _groups_meta(self):
        out = []
        for gi, g in enumerate(self.param_groups):
            out.append({k: v for k, v in g.items() if k != "params"} | {"step": self.step_count[gi]})
        return out


class MixedPrecisionOptimizer(MegatronOptimizer):
    """fp32 master + (optional) dynamic loss scaling over bf16/fp16 model params."""

    def __init__(self, param_groups, config: OptimizerConfig, grad_scaler: Optional[MegatronGradScaler], init_state_fn: Callable = lambda x: None):
        super().__init__(param_groups, config, init_state_fn)
        self.grad_scaler = grad_scaler
        if grad_scaler is None:
            assert not config.fp16, "fp16 expects a grad scaler"
        dev = torch.cuda.current_device() if torch.cuda.is_available() else "cpu"
        self.found_inf = torch.zeros(1, dtype=torch.float, device=dev)
        self._scale_one = torch.ones(1, dtype=torch.float, device=dev)
        self._clip_coeff = None

    def get_loss_scale(self):
        return self._scale_one if self.grad_scaler is None else self.grad_scaler.scale

    def _found_inf(self, norm: Optional[torch.Tensor]) -> bool:
        """fp16 only: a non-finite global norm ⇔ some grad overflowed (all-reduced already)."""
        if norm is None:
            norm = self.get_grad_norm()
        return not bool(torch.isfinite(norm))

    @torch.no_grad()
    def prepare_grads(self) -> bool:
        self._norm_cache = None
        if self.grad_scaler is not None:
            self._norm_cache = self.get_grad_norm()
            found = self._found_inf(self._norm_cache)
            self.grad_scaler.update(found)
            return found
        return False

    @torch.no_grad()
    def step_with_ready_grads(self) -> bool:
        scale = None
        if self.grad_scaler is not None:
            scale = self.grad_scaler.inv_scale.to(self.found_inf.device)
        if self._clip_coeff is not None:
            scale = self._clip_coeff.reshape(1) if scale is None else scale * self._clip_coeff
        self._apply_update(scale)
        return True

    @torch.no_grad()
    def step(self):
        timers = self.config.timers
        found_inf = self.prepare_grads()
        if found_inf:
            return False, None, None
        if timers is not None:
            timers("optimizer-clip-main-grad", log_level=1).start(barrier=self.config.barrier_with_L1_time)
        grad_norm = None
        self._clip_coeff = None
        if True:
            grad_norm = self._norm_cache if self._norm_cache is not None else self.get_grad_norm()
            if self.grad_scaler is not None:
                grad_norm = grad_norm * self.grad_scaler.inv_scale.to(grad_norm.device)[0]
            if self.config.clip_grad > 0.0:
                self._clip_coeff = clip_coefficient(grad_norm, self.config.clip_grad)
        if timers is not None:
            timers("optimizer-clip-main-grad").stop()
        num_zeros = self.count_zeros() if self.config.log_num_zeros_in_grad else None
        if timers is not None:
            timers("optimizer-inner-step", log_level=1).start(barrier=self.config.barrier_with_L1_time)
        ok = self.step_with_ready_grads()
        if timers is not None:
            timers("optimizer-inner-step").stop()
        return ok, grad_norm, num_zeros


class Float16OptimizerWithFloat16Params(MixedPrecisionOptimizer):
    """Whole-parameter slots: fp32 master per bf16/fp16 param; fp32 params update in place."""

    def __init__(self, param_groups, config: OptimizerConfig, grad_scaler: Optional[MegatronGradScaler] = None, init_state_fn: Callable = lambda x: None):
        super().__init__(param_groups, config, grad_scaler, init_state_fn)
        self.float16_groups, self.fp32_from_float16_groups, self.fp32_from_fp32_groups = [], [], []
        for gi, group in enumerate(self.param_groups):
            f16, f32m, f32 = [], [], []
            for p in group["params"]:
                if not p.requires_grad:
                    continue
                gf = (lambda _p=p: _p.main_grad if hasattr(_p, "main_grad") else _p.grad)
                if p.dtype in (torch.float16, torch.bfloat16):
                    master = p.detach().clone().float()
                    if hasattr(p, "shared"):
                        master.shared = p.shared
                    p.main_param = master
                    self.slots.append(Slot(p, gf, master, p.data, gi))
                    f16.append(p), f32m.append(master)
                elif p.dtype == torch.float32:
                    self.slots.append(Slot(p, gf, p.data, None, gi))
                    f32.append(p)
                else:
                    raise TypeError(f"unsupported parameter dtype {p.dtype}")
            self.float16_groups.append(f16), self.fp32_from_float16_groups.append(f32m), self.fp32_from_fp32_groups.append(f32)

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            _zero_grad_group(g["params"], set_to_none)

    def reload_model_params(self, state_dict=None):
        for s in self.slots:
            if s.lowp is not None:
                s.master.copy_(s.param.data)

    def state_dict(self):
        return {
            "optimizer": {"param_groups": self._groups_meta(), "state": {i: {k: v for k, v in self._slot_state(s).items() if k != "param"} for i, s in enumerate(self.slots)}},
            "grad_scaler": self.grad_scaler.state_dict() if self.grad_scaler else None,
            "fp32_from_fp16_params": [[s.master for s in self.slots if s.group == gi and s.lowp is not None] for gi in range(len(self.param_groups))],
        }

    def load_state_dict(self, state_dict):
        opt = state_dict["optimizer"]
        for gi, meta in enumerate(opt["param_groups"]):
            self.step_count[gi] = meta.get("step", 0)
            for k, v in meta.items():
                if k not in ("params", "step"):
                    self.param_groups[gi][k] = v
        for i, s in enumerate(self.slots):
            st = opt["state"].get(i) or opt["state"].get(str(i))
            if st:
                self._init_slot_state(s)
                for k, v in st.items():
                    if k.startswith("extra."):
                        name = k[len("extra."):]
                        if s.extra is None:
                            s.extra = {}
                        s.extra[name] = (v.clone() if v.dim() else (bool(v) if v.dtype == torch.bool else int(v))) if torch.is_tensor(v) else v
                    else:
                        getattr(s, k).copy_(v)
        if self.grad_scaler and state_dict.get("grad_scaler"):
            self.grad_scaler.load_state_dict(state_dict["grad_scaler"])
        masters = state_dict.get("fp32_from_fp16_params")
        if masters:
            for gi, lst in enumerate(masters):
                mine = [s for s in self.slots if s.group == gi and s.lowp is not None]
                for s, saved in zip(mine, lst):
                    s.master.copy_(saved)

    def sharded_state_dict(self, model_sharded_state_dict, is_loading: bool = False, metadata: Optional[dict] = None):
        """Optimizer tensors mirror the sharding of their model parameter."""
        from ..dist_checkpointing.optimizer import get_param_id_to_sharded_param_map, make_sharded_optimizer_tensor

        if is_loading:
            for s in self.slots:
                self._init_slot_state(s)
        id_map = get_param_id_to_sharded_param_map(model_sharded_state_dict, (s.param for s in self.slots))
        out = {"optimizer": {"param_groups": self._groups_meta(), "state": {}}, "grad_scaler": self.grad_scaler.state_dict() if self.grad_scaler else None}
        for i, s in enumerate(self.slots):
            model_sh = id_map[i]
            st = {}
            for k, v in self._slot_state(s).items():
                if v is None:
                    continue
                st[k if k != "param" else "fp32_param"] = make_sharded_optimizer_tensor(model_sh, v, prefix=f"optimizer.state.{k if k != 'param' else 'fp32_param'}")
            out["optimizer"]["state"][i] = st
        return out

    def load_sharded_state_dict(self, sd):
        for gi, meta in enumerate(sd["optimizer"]["param_groups"]):
            self.step_count[gi] = meta.get("step", 0)
        for i, s in enumerate(self.slots):
            st = sd["optimizer"]["state"][i]
            for k, v in st.items():
                dst = s.master if k == "fp32_param" else getattr(s, k)
                dst.copy_(v)
            if s.lowp is not None:
                s.lowp.copy_(s.master)


class FP32Optimizer(MegatronOptimizer):
    """All parameters already fp32: master is the parameter itself."""

    def __init__(self, param_groups, config: OptimizerConfig, init_state_fn: Callable = lambda x: None):
        super().__init__(param_groups, config, init_state_fn)
        dev = torch.cuda.current_device() if torch.cuda.is_available() else "cpu"
        self._scale = torch.tensor([1.0], dtype=torch.float, device=dev)
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.requires_grad:
                    self.slots.append(Slot(p, (lambda _p=p: _p.main_grad if hasattr(_p, "main_grad") else _p.grad), p.data, None, gi))

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            _zero_grad_group(g["params"], set_to_none)

    def get_loss_scale(self):
        return self._scale

    @torch.no_grad()
    def prepare_grads(self) -> bool:
        return False

    @torch.no_grad()
    def step_with_ready_grads(self) -> bool:
        self._apply_update(self._clip_coeff.reshape(1) if getattr(self, "_clip_coeff", None) is not None else None)
        return True

    @torch.no_grad()
    def step(self):
        self._clip_coeff = None
        grad_norm = self.get_grad_norm()
        if self.config.clip_grad > 0.0:
            self._clip_coeff = clip_coefficient(grad_norm, self.config.clip_grad)
        num_zeros = self.count_zeros() if self.config.log_num_zeros_in_grad else None
        ok = self.step_with_ready_grads()
        return ok, grad_norm, num_zeros

    state_dict = Float16OptimizerWithFloat16Params.state_dict
    load_state_dict = Float16OptimizerWithFloat16Params.load_state_dict
    sharded_state_dict = Float16OptimizerWithFloat16Params.sharded_state_dict
    load_sharded_state_dict = Float16OptimizerWithFloat16Params.load_sharded_state_dict
    grad_scaler = None


class ChainedOptimizer(MegatronOptimizer):
    """Several optimizers stepped as one (dense + expert params have different DP groups).
    The gradient norm is combined across members before any of them clips."""

    def __init__(self, chained_optimizers: List[MegatronOptimizer]):
        self.chained_optimizers = chained_optimizers
        self.config = chained_optimizers[0].config if chained_optimizers else None
        self.model_chunks = []
        for o in chained_optimizers:
            for c in getattr(o, "model_chunks", []):
                if c not in self.model_chunks:
                    self.model_chunks.append(c)
        self.is_stub_optimizer = all(getattr(o, "is_stub_optimizer", False) for o in chained_optimizers)

    @property
    def param_groups(self) -> List[dict]:
        return [g for o in self.chained_optimizers for g in o.param_groups]

    @property
    def slots(self):
        return [s for o in self.chained_optimizers for s in o.slots]

    def zero_grad(self, set_to_none=True):
        for o in self.chained_optimizers:
            o.zero_grad(set_to_none)

    def get_loss_scale(self):
        return self.chained_optimizers[0].get_loss_scale() if self.chained_optimizers else torch.tensor([1.0])

    def reload_model_params(self, state_dict=None):
        for o in self.chained_optimizers:
            o.reload_model_params(state_dict)

    def state_dict(self):
        return [o.state_dict() for o in self.chained_optimizers]

    def load_state_dict(self, state_dict):
        if isinstance(state_dict, dict):
            state_dict = [state_dict[k] for k in sorted(state_dict)]
        assert len(state_dict) == len(self.chained_optimizers)
        for o, sd in zip(self.chained_optimizers, state_dict):
            o.load_state_dict(sd)

    def sharded_state_dict(self, model_sharded_state_dict, is_loading=False, metadata=None, **kw):
        if len(self.chained_optimizers) == 1:
            return self.chained_optimizers[0].sharded_state_dict(model_sharded_state_dict, is_loading, metadata, **kw)
        out = {}
        for i, o in enumerate(self.chained_optimizers):
            sd = o.sharded_state_dict(model_sharded_state_dict, is_loading, metadata, **kw)
            from ..dist_checkpointing.utils import add_prefix_for_sharding

            add_prefix_for_sharding(sd, f"chained_{i}.")
            out[i] = sd
        return out

    def load_sharded_state_dict(self, sd):
        if len(self.chained_optimizers) == 1:
            return self.chained_optimizers[0].load_sharded_state_dict(sd)
        for i, o in enumerate(self.chained_optimizers):
            o.load_sharded_state_dict(sd[i])

    def get_grad_norm(self):
        sq = None
        for o in self.chained_optimizers:
            n = o.get_grad_norm().float() ** 2
            sq = n if sq is None else sq + n.to(sq.device)
        return sq.sqrt() if sq is not None else torch.zeros(())

    @torch.no_grad()
    def prepare_grads(self) -> bool:
        found = False
        for o in self.chained_optimizers:
            found |= o.prepare_grads()
        return found

    @torch.no_grad()
    def step_with_ready_grads(self) -> bool:
        ok = True
        for o in self.chained_optimizers:
            ok &= o.step_with_ready_grads()
        return ok

    @torch.no_grad()
    def step(self):
        if self.prepare_grads():
            return False, None, None
        grad_norm = self.get_grad_norm()
        for o in self.chained_optimizers:
            scaler = getattr(o, "grad_scaler", None)
            n = grad_norm * scaler.inv_scale.to(grad_norm.device)[0] if scaler is not None else grad_norm
            o._clip_coeff = clip_coefficient(n.to(o.slots[0].master.device) if o.slots else n, o.config.clip_grad) if o.config.clip_grad > 0.0 else None
        num_zeros = None
        if self.config is not None and self.config.log_num_zeros_in_grad:
            num_zeros = sum(o.count_zeros() for o in self.chained_optimizers)
        ok = self.step_with_ready_grads()
        for o in self.chained_optimizers:
            if hasattr(o, "_post_step"):
                o._post_step()
        return ok, grad_norm, num_zeros
