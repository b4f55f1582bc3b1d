"""Hang / fault detection for the training loop (reference ``training/ft_integration.py`` — section-based timeouts of
``nvidia_resiliency_ext.fault_tolerance``; that package is not a dependency here, so the monitor is self-contained).

The loop brackets its phases as *sections* (``setup``, ``step``, ``checkpointing``) and a watchdog thread checks that the open section
finishes within its timeout.  Timeouts start from configured values and, once ``calc_timeouts`` has seen enough samples, become
``safety_factor × max observed duration`` (persisted to ``ft_state.json`` next to the checkpoints, so a restarted job starts with learned
values).  On expiry the watchdog dumps every Python thread's stack, writes a ``hang_rank<r>.json`` record and — unless ``abort=False`` —
kills the process with SIGABRT so the launcher (torchrun ``--max-restarts`` / in-process restart) takes over.  A simulated fault can be
armed for tests (``maybe_setup_simulated_fault``)."""
from __future__ import annotations

import faulthandler
import json
import os
import random
import signal
import sys
import threading
import time
from typing import Dict, Optional

_STATE_FILE = "ft_state.json"


class FaultToleranceMonitor:
    def __init__(self, rank: int = 0, save_dir: Optional[str] = None, timeouts: Optional[Dict[str, float]] = None, safety_factor: float = 5.0,
                 min_samples: int = 16, poll_interval: float = 1.0, abort: bool = True, out_of_section_timeout: Optional[float] = None):
        self.rank, self.save_dir, self.abort = rank, save_dir, abort
        self.timeouts = dict(setup=1800.0, step=600.0, checkpointing=1800.0)
        self.timeouts.update(timeouts or {})
        self.out_of_section_timeout = out_of_section_timeout
        self.safety_factor, self.min_samples, self.poll = safety_factor, min_samples, poll_interval
        self.observed: Dict[str, list] = {}
        self._open: Optional[str] = None
        self._t_open = self._t_last = time.monotonic()
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.expired: Optional[dict] = None
        self.seen_checkpoints = False
        self._load_state()

    # ---- persistence ----
    def _state_path(self) -> Optional[str]:
        return os.path.join(self.save_dir, _STATE_FILE) if self.save_dir else None

    def _load_state(self) -> None:
        p = self._state_path()
        if p and os.path.exists(p):
            try:
                with open(p) as f:
                    self.timeouts.update(json.load(f).get("timeouts", {}))
            except (OSError, ValueError):
                pass

    def _save_state(self) -> None:
        p = self._state_path()
        if p and self.rank == 0:
            os.makedirs(self.save_dir, exist_ok=True)
            with open(p, "w") as f:
                json.dump({"timeouts": self.timeouts}, f)

    # ---- sections ----
    def start(self) -> "FaultToleranceMonitor":
        if self._thread is None:
            self._thread = threading.Thread(target=self._watch, name="ft-watchdog", daemon=True)
            self._thread.start()
        return self

    def shutdown(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2 * self.poll + 1)
            self._thread = None

    def start_section(self, name: str) -> None:
        with self._lock:
            self._open, self._t_open = name, time.monotonic()

    def end_section(self, name: str) -> None:
        with self._lock:
            if self._open == name:
                self.observed.setdefault(name, []).append(time.monotonic() - self._t_open)
                self._open = None
                self._t_last = time.monotonic()

    def calc_timeouts(self) -> Dict[str, float]:
        """Replace configured timeouts by learned ones for every section with enough samples."""
        for name, xs in self.observed.items():
            if len(xs) >= (1 if name != "step" else self.min_samples):
                self.timeouts[name] = max(self.safety_factor * max(xs), 1.0)
        self._save_state()
        return dict(self.timeouts)

    # ---- watchdog ----
    def _watch(self) -> None:
        while not self._stop.wait(self.poll):
            with self._lock:
                name, t0, t_last = self._open, self._t_open, self._t_last
            now = time.monotonic()
            if name is not None and now - t0 > self.timeouts.get(name, float("inf")):
                self._expire(name, now - t0)
                return
            if name is None and self.out_of_section_timeout and now - t_last > self.out_of_section_timeout:
                self._expire("out_of_section", now - t_last)
                return

    def _expire(self, name: str, elapsed: float) -> None:
        self.expired = {"rank": self.rank, "section": name, "elapsed_s": round(elapsed, 2), "timeout_s": self.timeouts.get(name, self.out_of_section_timeout),
                        "time": time.time()}
        sys.stderr.write(f"[ft] rank {self.rank}: section '{name}' exceeded its timeout ({elapsed:.1f}s); dumping stacks\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        if self.save_dir:
            try:
                os.makedirs(self.save_dir, exist_ok=True)
                with open(os.path.join(self.save_dir, f"hang_rank{self.rank}.json"), "w") as f:
                    json.dump(self.expired, f)
            except OSError:
                pass
        if self.abort:
            os.kill(os.getpid(), signal.SIGABRT)


_MON: Optional[FaultToleranceMonitor] = None


def setup(args, rank: int = 0) -> FaultToleranceMonitor:
    global _MON
    extra = {}
    if getattr(args, "ft_num_warmup_iters", None):          # reference --ft-num-warmup-iters: step samples needed before the learned step timeout replaces the configured one
        extra["min_samples"] = int(args.ft_num_warmup_iters)
    _MON = FaultToleranceMonitor(rank=rank, save_dir=getattr(args, "save", None), **extra,
                                 timeouts={k: v for k, v in dict(setup=getattr(args, "ft_timeout_setup", None), step=getattr(args, "ft_timeout_step", None),
                                                                 checkpointing=getattr(args, "ft_timeout_checkpointing", None)).items() if v}).start()
    _MON.start_section("setup")
    return _MON


def get_monitor() -> Optional[FaultToleranceMonitor]:
    return _MON


def on_training_step_start() -> None:
    if _MON is not None:
        if _MON._open == "setup":
            _MON.end_section("setup")
        _MON.start_section("step")


def on_training_step_end() -> None:
    if _MON is not None:
        _MON.end_section("step")


def on_eval_step_start() -> None:
    on_training_step_start()


def on_eval_step_end() -> None:
    on_training_step_end()


def on_checkpointing_start() -> None:
    if _MON is not None:
        _MON.start_section("checkpointing")


def on_checkpointing_end(is_async_finalization: bool = False) -> None:
    if _MON is not None:
        _MON.end_section("checkpointing")
        _MON.seen_checkpoints = True
        _MON.calc_timeouts()


def on_checkpoint_loaded(is_local_chkpt: bool = False) -> None:
    if _MON is not None and _MON._open == "setup":
        pass  # still inside setup: nothing to time separately


def shutdown() -> None:
    global _MON
    if _MON is not None:
        _MON.shutdown()
        _MON = None


def maybe_setup_simulated_fault(args=None, rank: int = 0, world: int = 1, kind: Optional[str] = None, delay_s: Optional[float] = None, target_rank: Optional[int] = None,
                                seed: int = 0) -> Optional[threading.Thread]:
    """Arm a fault for resiliency tests: after ``delay_s`` the chosen rank either gets SIGKILL (``rank_killed``) or stops making progress
    (``rank_hung``: SIGSTOP).  Spec can come from ``args.simulated_fault`` = "kind:delay[:rank]"."""
    spec = kind or (getattr(args, "simulated_fault", None) if args is not None else None)
    if not spec:
        return None
    if ":" in spec:
        parts = spec.split(":")
        spec, delay_s = parts[0], float(parts[1])
        if len(parts) > 2:
            target_rank = int(parts[2])
    if target_rank is None:
        target_rank = random.Random(seed).randrange(world)
    if rank != target_rank:
        return None
    sig = {"rank_killed": signal.SIGKILL, "rank_hung": signal.SIGSTOP}[spec]

    def fire():
        time.sleep(delay_s or 0.0)
        os.kill(os.getpid(), sig)

    t = threading.Thread(target=fire, name="ft-simulated-fault", daemon=True)
    t.start()
    return t
