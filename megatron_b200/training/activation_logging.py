"""Periodic activation / gradient statistics dumps (reference ``training/activation_logging.py``, ``dgrad_logging.py``; wired at
``training.py:3024-3036``).  Forward hooks record output statistics of the selected modules, full-backward hooks record the statistics of
the gradient w.r.t. their outputs (dgrad); both are only armed for the iterations the schedule selects, so the steady state has no hooks
installed and no host syncs.  Records are written as one JSON file per (iteration, rank)."""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional

import torch


def _stats(t: torch.Tensor) -> Dict[str, float]:
    f = t.detach().float()
    return {"shape": list(t.shape), "norm": f.norm().item(), "absmax": f.abs().max().item() if f.numel() else 0.0, "mean": f.mean().item() if f.numel() else 0.0,
            "std": f.std().item() if f.numel() > 1 else 0.0, "nan": int(torch.isnan(f).sum().item()), "inf": int(torch.isinf(f).sum().item())}


def _first_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (tuple, list)):
        for e in x:
            t = _first_tensor(e)
            if t is not None:
                return t
    return None


class _StatLogger:
    kind = "activation"

    def __init__(self, model, out_dir: str, interval: int, pattern: str = r".*(self_attention|mlp|input_layernorm|pre_mlp_layernorm)$", rank: int = 0,
                 save_tensors: bool = False):
        self.models = model if isinstance(model, (list, tuple)) else [model]
        self.out_dir, self.interval, self.rank, self.save_tensors = out_dir, max(int(interval), 0), rank, save_tensors
        self.rx = re.compile(pattern)
        self.handles: List = []
        self.records: Dict[str, dict] = {}
        self.iteration = 0

    def _selected(self):
        for ci, m in enumerate(self.models):
            for name, mod in m.named_modules():
                if name and self.rx.match(name):
                    yield f"chunk{ci}.{name}", mod

    def _install(self) -> None:
        raise NotImplementedError

    def begin_iteration(self, iteration: int) -> bool:
        """Arm the hooks when ``iteration`` is a logging iteration; returns whether it is."""
        self.iteration = iteration
        if not self.interval or iteration % self.interval:
            return False
        self.records = {}
        self._install()
        return True

    def end_iteration(self) -> Optional[str]:
        if not self.handles:
            return None
        for h in self.handles:
            h.remove()
        self.handles = []
        os.makedirs(self.out_dir, exist_ok=True)
        path = os.path.join(self.out_dir, f"{self.kind}_iter{self.iteration:07d}_rank{self.rank:05d}.json")
        with open(path, "w") as f:
            json.dump(self.records, f, indent=1)
        return path

    def _record(self, name: str, t: Optional[torch.Tensor]) -> None:
        if t is None:
            return
        key = name if name not in self.records else f"{name}#{sum(k.startswith(name) for k in self.records)}"   # micro-batches append
        self.records[key] = _stats(t)
        if self.save_tensors:
            os.makedirs(self.out_dir, exist_ok=True)
            torch.save(t.detach().cpu(), os.path.join(self.out_dir, f"{self.kind}_iter{self.iteration:07d}_rank{self.rank:05d}_{key.replace('.', '_')}.pt"))


class ActivationLogger(_StatLogger):
    kind = "activation"

    def _install(self) -> None:
        for name, mod in self._selected():
            self.handles.append(mod.register_forward_hook(lambda m, i, o, _n=name: self._record(_n, _first_tensor(o))))


class DgradLogger(_StatLogger):
    kind = "dgrad"

    def _install(self) -> None:
        for name, mod in self._selected():
            self.handles.append(mod.register_full_backward_hook(lambda m, gi, go, _n=name: self._record(_n, _first_tensor(go))))


class WgradLogger(_StatLogger):
    """Statistics of the parameter gradients (main_grad when the DDP buffers own them) after backward."""

    kind = "wgrad"

    def _install(self) -> None:
        self.handles.append(type("_H", (), {"remove": lambda s: None})())

    def collect(self) -> None:
        for ci, m in enumerate(self.models):
            for name, p in m.named_parameters():
                g = getattr(p, "main_grad", None)
                g = g if g is not None else p.grad
                if g is not None:
                    self._record(f"chunk{ci}.{name}", g)
