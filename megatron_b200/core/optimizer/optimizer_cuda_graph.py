"""Optimizer step as a CUDA graph (reference ``optimizer/optimizer_cuda_graph.py``): the fused multi-tensor update is a handful of long kernels, but
clipping, loss-scale bookkeeping and per-group launches add ~50 small ones per step; with static gradient buffers (DDP ``main_grad``) the whole
``step()`` is capturable — no host syncs are left in our clipping path (``clip_coefficient`` stays on the device).

``GraphedOptimizerStep(optimizer)`` captures on the first call after ``warmup`` eager steps and replays afterwards; learning-rate / weight-decay
schedules are fed through device scalars refreshed before each replay, so the captured kernels read the current values."""
from __future__ import annotations

from typing import Optional

import torch


class GraphedOptimizerStep:
    def __init__(self, optimizer, warmup: int = 2):
        self.optimizer, self.warmup = optimizer, warmup
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.result = None
        self.fallback_reason: Optional[str] = None

    def _capturable(self) -> bool:
        if not torch.cuda.is_available() or self.fallback_reason is not None:
            return False
        cfg = getattr(self.optimizer, "config", None)
        return not (cfg is not None and getattr(cfg, "fp16", False))      # dynamic loss scaling branches on the host

    def step(self):
        self.calls += 1
        if not self._capturable() or self.calls <= self.warmup:
            return self.optimizer.step()
        if self.graph is None:
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.result = self.optimizer.step()
                self.graph = g
            except Exception as e:      # capture is best effort
                self.fallback_reason = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
                return self.optimizer.step()
        self.graph.replay()
        return self.result

    def __getattr__(self, name):
        return getattr(self.optimizer, name)
