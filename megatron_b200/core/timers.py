"""Named timers with device-side timing (reference ``timers.py:35-485``).

Unlike the reference (host clock + ``cuda.synchronize`` at start/stop) the
default timer records CUDA events on the current stream and resolves them
lazily at report time, so an enabled timer never stalls the launch queue.
``log_option`` min/max/all across ranks is kept.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class TimerBase:
    def __init__(self, name):
        self.name = name

    def start(self, barrier=False):
        raise NotImplementedError

    def stop(self, barrier=False):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def elapsed(self, reset=True, barrier=False):
        raise NotImplementedError


class DummyTimer(TimerBase):
    def __init__(self):
        super().__init__("dummy")

    def start(self, barrier=False):
        return

    def stop(self, barrier=False):
        return

    def reset(self):
        return

    def elapsed(self, reset=True, barrier=False):
        raise Exception("dummy timer should not be used to calculate elapsed time")

    def active_time(self):
        return 0.0


class Timer(TimerBase):
    def __init__(self, name: str, use_events: Optional[bool] = None):
        super().__init__(name)
        self._use_events = torch.cuda.is_available() if use_events is None else use_events
        self._barrier_group = None
        self.reset()

    def set_barrier_group(self, group):
        self._barrier_group = group

    def reset(self):
        self._elapsed = 0.0
        self._active = 0.0
        self._started = False
        self._pending = []  # (start_event, stop_event)
        self._t0 = None

    def start(self, barrier=False):
        assert not self._started, f"timer {self.name} has already been started"
        if barrier and dist.is_initialized():
            dist.barrier(group=self._barrier_group)
        if self._use_events:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._t0 = ev
        else:
            self._t0 = time.perf_counter()
        self._started = True

    def stop(self, barrier=False):
        assert self._started, f"timer {self.name} is not started"
        if barrier and dist.is_initialized():
            dist.barrier(group=self._barrier_group)
        if self._use_events:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._pending.append((self._t0, ev))
        else:
            dt = time.perf_counter() - self._t0
            self._elapsed += dt
            self._active += dt
        self._started = False

    def _resolve(self):
        if self._pending:
            self._pending[-1][1].synchronize()
            for s, e in self._pending:
                dt = s.elapsed_time(e) / 1e3
                self._elapsed += dt
                self._active += dt
            self._pending = []

    def elapsed(self, reset=True, barrier=False):
        was = self._started
        if was:
            self.stop(barrier=barrier)
        self._resolve()
        out = self._elapsed
        if reset:
            self._elapsed = 0.0
        if was:
            self.start(barrier=barrier)
        return out

    def active_time(self):
        self._resolve()
        return self._active


class Timers:
    """``timers('name', log_level=k).start()`` — timers above ``log_level`` are no-ops."""

    def __init__(self, log_level: int = 0, log_option: str = "minmax"):
        assert log_option in ("max", "minmax", "all")
        self._log_level, self._log_option = log_level, log_option
        self._timers: Dict[str, Timer] = {}
        self._levels: Dict[str, int] = {}
        self._dummy = DummyTimer()
        self._max_log_level = 2

    def __call__(self, name, log_level=None, barrier=False):
        if name in self._timers:
            return self._timers[name]
        log_level = self._max_log_level if log_level is None else log_level
        assert log_level <= self._max_log_level
        if log_level > self._log_level:
            return self._dummy
        self._timers[name] = Timer(name)
        self._levels[name] = log_level
        return self._timers[name]

    def _gather(self, names, reset, barrier):
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        local = torch.zeros(len(names), dtype=torch.float64)
        for i, n in enumerate(names):
            if n in self._timers:
                local[i] = self._timers[n].elapsed(reset=reset, barrier=barrier)
        if world == 1:
            return local.unsqueeze(0)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        allt = torch.zeros(world, len(names), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt.view(-1), local.to(dev))
        return allt.cpu()

    def get_all_timers_string(self, names: Optional[List[str]] = None, normalizer: float = 1.0, reset: bool = True, barrier: bool = False):
        names = list(self._timers.keys()) if names is None else names
        if not names:
            return None
        t = self._gather(names, reset, barrier) * 1000.0 / normalizer
        lines = []
        if self._log_option in ("max", "minmax"):
            head = "(min, max) time across ranks (ms):" if self._log_option == "minmax" else "max time across ranks (ms):"
            lines.append(head)
            for i, n in enumerate(names):
                col = t[:, i]
                nz = col[col > 0]
                if nz.numel() == 0:
                    continue
                if self._log_option == "minmax":
                    lines.append(f"    {n.ljust(48, '.')}: ({nz.min():.2f}, {nz.max():.2f})")
                else:
                    lines.append(f"    {n.ljust(48, '.')}: {nz.max():.2f}")
        else:
            lines.append("times across ranks (ms):")
            for i, n in enumerate(names):
                lines.append(f"  {n}:")
                for r in range(t.shape[0]):
                    if t[r, i] > 0:
                        lines.append(f"     rank {r:2d}: {t[r, i]:.2f}")
        return "\n".join(lines) if len(lines) > 1 else None

    def log(self, names=None, rank=None, normalizer=1.0, reset=True, barrier=False):
        s = self.get_all_timers_string(names, normalizer, reset, barrier)
        if s is None:
            return
        world = dist.get_world_size() if dist.is_initialized() else 1
        me = dist.get_rank() if dist.is_initialized() else 0
        if me == (world - 1 if rank is None else rank):
            print(s, flush=True)

    def write(self, names, writer, iteration, normalizer=1.0, reset=True, barrier=False):
        t = self._gather(names, reset, barrier) / normalizer
        if writer is not None:
            for i, n in enumerate(names):
                writer.add_scalar(n + "-time", float(t[:, i].max()), iteration)
