"""Optimizer with part of its state on the host (reference ``optimizer/cpu_offloading/hybrid_optimizer.py`` — ``HybridDeviceOptimizer``).

``offload_fraction`` of the parameter elements (largest parameters first, so few big copies) keep their fp32 copy and optimizer state in PINNED host
memory and are stepped by a CPU optimizer; the rest is stepped on the GPU as usual.  Per step:

    D2H  gradients of the offloaded parameters   (side stream, overlapped with the GPU sub-step)
    CPU  optimizer step on the host copies
    H2D  updated parameters back into the model  (side stream; the caller's next forward waits on ``self.h2d_event``)

On a B200 node the trade is 180 GB of HBM against a ~55 GB/s PCIe Gen5 link per GPU: offloading the 12 bytes/param of Adam state of an 8B model
moves 16 GB of gradients and 16 GB of parameters per step (≈ 0.6 s), so this is a capacity tool for models that otherwise do not fit, not a speed-up.
Works with any ``torch.optim`` class pair; state dicts are merged so checkpoints do not depend on the offload fraction."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Type

import torch


class HybridDeviceOptimizer(torch.optim.Optimizer):
    def __init__(self, params: Iterable, offload_fraction: float = 1.0, cpu_optimizer_cls: Type[torch.optim.Optimizer] = torch.optim.AdamW,
                 gpu_optimizer_cls: Optional[Type[torch.optim.Optimizer]] = None, pin_cpu_params: bool = True, pin_cpu_grads: bool = True,
                 overlap_cpu_optimizer_d2h_h2d: bool = True, **defaults):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        super().__init__(groups, defaults)
        self.offload_fraction = float(offload_fraction)
        gpu_optimizer_cls = gpu_optimizer_cls or cpu_optimizer_cls
        self.pin = torch.cuda.is_available()
        self.overlap = overlap_cpu_optimizer_d2h_h2d and torch.cuda.is_available()
        self._d2h = torch.cuda.Stream() if self.overlap else None
        self._h2d = torch.cuda.Stream() if self.overlap else None
        self.h2d_event = None
        # ---- choose what to offload: biggest parameters first until the element budget is met ----
        allp = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        budget = self.offload_fraction * sum(p.numel() for p in allp)
        self.offloaded: Dict[torch.nn.Parameter, torch.Tensor] = {}
        acc = 0
        for p in sorted(allp, key=lambda t: -t.numel()):
            if acc >= budget or budget <= 0:
                break
            host = torch.empty(p.shape, dtype=torch.float32, device="cpu", pin_memory=self.pin and pin_cpu_params)
            host.copy_(p.detach())
            host.grad = torch.zeros(p.shape, dtype=torch.float32, device="cpu", pin_memory=self.pin and pin_cpu_grads)
            self.offloaded[p] = host
            acc += p.numel()
        cpu_groups, gpu_groups = [], []
        for g in self.param_groups:
            meta = {k: v for k, v in g.items() if k != "params"}
            c = [self.offloaded[p] for p in g["params"] if p in self.offloaded]
            d = [p for p in g["params"] if p not in self.offloaded and p.requires_grad]
            cpu_groups.append(dict(meta, params=c))
            gpu_groups.append(dict(meta, params=d))
        self.cpu_optimizer = cpu_optimizer_cls([g for g in cpu_groups if g["params"]], **defaults) if any(g["params"] for g in cpu_groups) else None
        self.gpu_optimizer = gpu_optimizer_cls([g for g in gpu_groups if g["params"]], **defaults) if any(g["params"] for g in gpu_groups) else None
        self._cpu_groups, self._gpu_groups = cpu_groups, gpu_groups

    # the lr / wd schedulers of the training loop mutate ``self.param_groups``: mirror those fields into the sub-optimizers before stepping
    def _sync_hyperparams(self) -> None:
        for opt, groups in ((self.cpu_optimizer, self._cpu_groups), (self.gpu_optimizer, self._gpu_groups)):
            if opt is None:
                continue
            live = [g for g in groups if g["params"]]
            for sub, mine in zip(opt.param_groups, live):
                src = self.param_groups[groups.index(mine)]
                for k, v in src.items():
                    if k != "params":
                        sub[k] = v

    @torch.no_grad()
    def step(self, closure=None):
        self._sync_hyperparams()
        cur = torch.cuda.current_stream() if self.overlap else None
        if self.overlap:
            self._d2h.wait_stream(cur)
        for p, host in self.offloaded.items():
            g = p.main_grad if hasattr(p, "main_grad") else p.grad
            if g is None:
                host.grad.zero_()
                continue
            if self.overlap:
                with torch.cuda.stream(self._d2h):
                    host.grad.copy_(g, non_blocking=True)
            else:
                host.grad.copy_(g)
        if self.gpu_optimizer is not None:
            for grp in self.gpu_optimizer.param_groups:          # feed main_grad to plain torch optimizers
                for p in grp["params"]:
                    if hasattr(p, "main_grad") and p.grad is None:
                        p.grad = p.main_grad.to(p.dtype)
            self.gpu_optimizer.step()                            # runs while the gradients of the offloaded part travel
        if self.overlap:
            self._d2h.synchronize()
        if self.cpu_optimizer is not None:
            self.cpu_optimizer.step()
        for p, host in self.offloaded.items():
            if self.overlap:
                with torch.cuda.stream(self._h2d):
                    p.data.copy_(host, non_blocking=True)
            else:
                p.data.copy_(host)
        if self.overlap:
            self.h2d_event = torch.cuda.Event()
            self.h2d_event.record(self._h2d)
            cur.wait_event(self.h2d_event)

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    # ---- checkpoint: keyed by the model parameter order, independent of where the state lives ----
    def state_dict(self):
        out = {"offload_fraction": self.offload_fraction, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups], "state": {}}
        idx = 0
        cpu_state = self.cpu_optimizer.state if self.cpu_optimizer is not None else {}
        gpu_state = self.gpu_optimizer.state if self.gpu_optimizer is not None else {}
        for g in self.param_groups:
            for p in g["params"]:
                st = cpu_state.get(self.offloaded[p]) if p in self.offloaded else gpu_state.get(p)
                if st:
                    out["state"][idx] = {k: (v.detach().cpu().clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
                idx += 1
        return out

    def load_state_dict(self, sd):
        for g, saved in zip(self.param_groups, sd.get("param_groups", [])):      # hyper-parameters (lr after decay, ...) travel with the checkpoint
            g.update({k: v for k, v in saved.items() if k != "params"})
        idx = 0
        for g in self.param_groups:
            for p in g["params"]:
                st = sd["state"].get(idx)
                idx += 1
                if st is None:
                    continue
                if p in self.offloaded:
                    tgt, key = self.cpu_optimizer.state, self.offloaded[p]
                else:
                    tgt, key = self.gpu_optimizer.state, p
                tgt[key] = {k: (v.to(key.device) if isinstance(v, torch.Tensor) and v.dim() > 0 else v) for k, v in st.items()}
