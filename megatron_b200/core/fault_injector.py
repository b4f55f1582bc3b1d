"""Fault injection for resiliency testing (reference ``core/fault_injector.py:48-233``, which delegates to nvidia-resiliency-ext).

Self-contained: a daemon thread on the selected rank(s) fires ONE fault after ``delay`` seconds or at a given training iteration.
Fault kinds: ``gpu_sleep`` (spin kernel → straggler), ``gpu_error`` (illegal memory access → sticky CUDA error), ``workload_exc``
(Python exception in the training thread at the next ``maybe_raise()``), ``sigkill`` / ``sigterm`` / ``sigstop`` (signals to self),
``os_abort``.  Used by the rerun-state-machine / checkpoint-resume tests to prove recovery paths."""
from __future__ import annotations

import enum
import os
import random
import signal
import threading
import time
from dataclasses import dataclass
from typing import Optional, Sequence

import torch


class Fault(str, enum.Enum):
    GPU_SLEEP = "gpu_sleep"
    GPU_ERROR = "gpu_error"
    WORKLOAD_EXC = "workload_exc"
    SIGKILL = "sigkill"
    SIGTERM = "sigterm"
    SIGSTOP = "sigstop"
    OS_ABORT = "os_abort"


class InjectedFaultError(RuntimeError):
    pass


@dataclass
class FaultInjectorConfig:
    fault_type: Fault = Fault.WORKLOAD_EXC
    ranks: Optional[Sequence[int]] = None     # None → one random rank (same choice on every rank via the seed)
    delay_s: Optional[float] = None           # fire after this many seconds …
    at_iteration: Optional[int] = None        # … or when the training loop reports this iteration
    gpu_sleep_s: float = 30.0
    seed: int = 1234


class FaultInjector:
    def __init__(self, cfg: FaultInjectorConfig, rank: Optional[int] = None, world_size: Optional[int] = None):
        import torch.distributed as dist

        self.cfg = cfg
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        ranks = list(cfg.ranks) if cfg.ranks is not None else [random.Random(cfg.seed).randrange(world)]
        self.armed = self.rank in ranks
        self.fired = False
        self._pending_exc = False
        self._thread: Optional[threading.Thread] = None
        if self.armed and cfg.delay_s is not None:
            self._thread = threading.Thread(target=self._timer, daemon=True)
            self._thread.start()

    def _timer(self):
        time.sleep(self.cfg.delay_s)
        self.fire()

    def on_iteration(self, iteration: int):
        """Call once per training iteration (``training.train`` does when ``--inject-fault`` is set)."""
        if self.armed and not self.fired and self.cfg.at_iteration is not None and iteration >= self.cfg.at_iteration:
            self.fire()
        self.maybe_raise()

    def maybe_raise(self):
        if self._pending_exc:
            self._pending_exc = False
            raise InjectedFaultError(f"injected workload exception on rank {self.rank}")

    def fire(self):
        if self.fired:
            return
        self.fired = True
        k = Fault(self.cfg.fault_type)
        if k == Fault.WORKLOAD_EXC:
            self._pending_exc = True
        elif k == Fault.GPU_SLEEP:
            if torch.cuda.is_available():
                torch.cuda._sleep(int(self.cfg.gpu_sleep_s * 1.5e9))
            else:
                time.sleep(self.cfg.gpu_sleep_s)
        elif k == Fault.GPU_ERROR:
            if torch.cuda.is_available():
                bad = torch.empty(1, device="cuda")
                idx = torch.tensor([1 << 40], device="cuda")
                bad[idx] = 1.0  # out-of-bounds device write → device-side assert / illegal address
                torch.cuda.synchronize()
            else:
                self._pending_exc = True
        elif k == Fault.SIGKILL:
            os.kill(os.getpid(), signal.SIGKILL)
        elif k == Fault.SIGTERM:
            os.kill(os.getpid(), signal.SIGTERM)
        elif k == Fault.SIGSTOP:
            os.kill(os.getpid(), signal.SIGSTOP)
        elif k == Fault.OS_ABORT:
            os.abort()
