"""Learning-rate and weight-decay schedules (reference ``optimizer_param_scheduler.py:100-404``)."""
from __future__ import annotations

import math
from typing import Optional


class OptimizerParamScheduler:
    """Warm-up then {constant|linear|cosine|inverse-square-root|WSD} decay; wd may ramp too.
    ``step(increment)`` is called with the number of *samples* consumed."""

    def __init__(self, optimizer, init_lr: float, max_lr: float, min_lr: float, lr_warmup_steps: int, lr_decay_steps: int,
                 lr_decay_style: str, start_wd: float, end_wd: float, wd_incr_steps: int, wd_incr_style: str,
                 use_checkpoint_opt_param_scheduler: bool = True, override_opt_param_scheduler: bool = False,
                 wsd_decay_steps: Optional[int] = None, lr_wsd_decay_style: Optional[str] = None):
        self.optimizer = optimizer
        self.init_lr, self.max_lr, self.min_lr = init_lr, float(max_lr), min_lr
        assert 0.0 <= self.min_lr <= self.max_lr and self.init_lr <= self.max_lr
        self.lr_warmup_steps, self.lr_decay_steps = lr_warmup_steps, lr_decay_steps
        assert self.lr_decay_steps > 0 and self.lr_warmup_steps < self.lr_decay_steps
        self.lr_decay_style = lr_decay_style
        self.wsd_decay_steps, self.lr_wsd_decay_style = wsd_decay_steps, lr_wsd_decay_style
        if lr_decay_style == "WSD":
            assert wsd_decay_steps is not None
        self.start_wd, self.end_wd = start_wd, end_wd
        assert 0.0 <= start_wd <= end_wd
        self.wd_incr_steps, self.wd_incr_style = wd_incr_steps, wd_incr_style
        self.override_opt_param_scheduler = override_opt_param_scheduler
        self.use_checkpoint_opt_param_scheduler = use_checkpoint_opt_param_scheduler
        assert not (override_opt_param_scheduler and use_checkpoint_opt_param_scheduler), "both override and use-checkpoint are set"
        self.num_steps = 0
        self.step(0)

    def get_wd(self, param_group: Optional[dict] = None) -> float:
        """``param_group`` may carry its own ``start_wd`` / ``end_wd`` (reference ``ParamGroupOverride``)."""
        start_wd = self.start_wd if param_group is None else param_group.get("start_wd", self.start_wd)
        end_wd = self.end_wd if param_group is None else param_group.get("end_wd", self.end_wd)
        if self.num_steps > self.wd_incr_steps:
            return end_wd
        if self.wd_incr_style == "constant":
            assert start_wd == end_wd
            return end_wd
        r = float(self.num_steps) / float(self.wd_incr_steps)
        if self.wd_incr_style == "linear":
            c = r
        elif self.wd_incr_style == "cosine":
            c = 0.5 * (math.cos(math.pi * (1 - r)) + 1.0)
        else:
            raise Exception(f"{self.wd_incr_style} weight decay increment style is not supported")
        return start_wd + c * (end_wd - start_wd)

    def get_lr(self, param_group: dict) -> float:
        max_lr = param_group.get("max_lr", self.max_lr)
        min_lr = param_group.get("min_lr", self.min_lr)
        if self.lr_warmup_steps > 0 and self.num_steps <= self.lr_warmup_steps:
            return self.init_lr + (max_lr - self.init_lr) * float(self.num_steps) / float(self.lr_warmup_steps)
        if self.lr_decay_style == "constant":
            return max_lr
        if self.num_steps > self.lr_decay_steps:
            return min_lr
        if self.lr_decay_style == "inverse-square-root":
            w = max(self.lr_warmup_steps, 1)
            return max(min_lr, max_lr * w**0.5 / max(self.num_steps, 1) ** 0.5)
        n, d = self.num_steps - self.lr_warmup_steps, self.lr_decay_steps - self.lr_warmup_steps
        r = float(n) / float(d)
        delta = max_lr - min_lr
        if self.lr_decay_style == "linear":
            c = 1.0 - r
        elif self.lr_decay_style == "cosine":
            c = 0.5 * (math.cos(math.pi * r) + 1.0)
        elif self.lr_decay_style == "WSD":
            anneal_start = self.lr_decay_steps - self.wsd_decay_steps
            if self.num_steps <= anneal_start:
                c = 1.0
            else:
                rr = float(self.num_steps - anneal_start) / float(self.wsd_decay_steps)
                style = self.lr_wsd_decay_style
                c = {"linear": 1.0 - rr, "cosine": 0.5 * (math.cos(math.pi * rr) + 1.0), "exponential": 2.0 * (0.5**rr) - 1.0, "minus_sqrt": 1.0 - math.sqrt(rr)}[style]
        else:
            raise Exception(f"{self.lr_decay_style} decay style is not supported")
        return min_lr + c * delta

    def step(self, increment: int) -> None:
        self.num_steps += increment
        for g in self.optimizer.param_groups:
            lr = self.get_lr(g) * g.get("lr_mult", 1.0)      # lr_mult: this framework's multiplier; the reference folds it into per-group max_lr / min_lr
            if hasattr(g.get("lr"), "fill_"):
                g["lr"].fill_(lr)                              # tensor lr (captured optimizers): update in place
            else:
                g["lr"] = lr
            g["weight_decay"] = self.get_wd(g) * g.get("wd_mult", 1.0)

    def state_dict(self) -> dict:
        return {k: getattr(self, k) for k in ("max_lr", "lr_warmup_steps", "num_steps", "lr_decay_style", "lr_decay_steps", "min_lr", "start_wd", "end_wd", "wd_incr_style", "wd_incr_steps")}

    def _check_and_set(self, cls_value, sd_value, name):
        if self.override_opt_param_scheduler:
            return cls_value
        if not self.use_checkpoint_opt_param_scheduler:
            assert cls_value == sd_value, f"OptimizerParamScheduler: class input value {cls_value} and checkpoint value {sd_value} for {name} do not match"
        return sd_value

    def load_state_dict(self, sd: dict) -> None:
        for name in ("max_lr", "min_lr", "lr_warmup_steps", "lr_decay_steps", "lr_decay_style", "start_wd", "end_wd", "wd_incr_steps", "wd_incr_style"):
            if name in sd:
                setattr(self, name, self._check_and_set(getattr(self, name), sd[name], name))
        self.num_steps = 0
        self.step(sd.get("num_steps", sd.get("num_iters", 0)))
