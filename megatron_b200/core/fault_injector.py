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
    start_iteration: Optional[int] = None     # with ``delay_s``: the countdown starts when the training loop reports this iteration, not at construction

    @classmethod
    def from_args(cls, args, world_size: int) -> Optional["FaultInjectorConfig"]:
        """The reference command line (``--fault-injector-*``, ``training/config/resilience_config.py:FaultInjectorConfig``): explicit ``ranks`` or
        ``num_ranks`` seeded picks; one fault kind drawn from ``fault_types`` by ``fault_probabilities``; a fixed ``fault_delay`` or an exponential draw with
        mean ``mtti_seconds`` (plus ``offset_seconds``), counted from ``delay_start_iteration`` when given.  ``None`` when no fault is configured."""
        names = ("fault_injector_ranks", "fault_injector_num_ranks", "fault_injector_fault_types", "fault_injector_fault_probabilities", "fault_injector_fault_delay",
                 "fault_injector_delay_start_iteration", "fault_injector_mtti_seconds", "fault_injector_offset_seconds", "fault_injector_seed")
        vals = {n[len("fault_injector_"):]: getattr(args, n, None) for n in names}
        g = vals.get
        kinds = [k.strip().lower() for k in (g("fault_types") or "").split(",") if k.strip()]
        if not kinds:
            return None
        rng = random.Random(g("seed") if g("seed") is not None else 1234)
        ranks = [int(r) for r in str(g("ranks")).split(",")] if g("ranks") else sorted(rng.sample(range(world_size), min(g("num_ranks") or 1, world_size)))
        probs = [float(x) for x in (g("fault_probabilities") or "").split(",") if x.strip()] or [1.0] * len(kinds)
        assert len(probs) == len(kinds), "--fault-injector-fault-probabilities needs one entry per fault type"
        kind = rng.choices(kinds, weights=probs)[0]
        delay = g("fault_delay")
        if delay is None and g("mtti_seconds"):
            delay = rng.expovariate(1.0 / g("mtti_seconds"))
        delay = (delay or 0.0) + (g("offset_seconds") or 0.0)
        return cls(fault_type=Fault(kind), ranks=ranks, delay_s=delay, seed=g("seed") if g("seed") is not None else 1234, start_iteration=g("delay_start_iteration"))


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
        if self.armed and cfg.delay_s is not None and cfg.start_iteration is None:
            self._start_timer()

    def _start_timer(self):
        if self._thread is None:
            self._thread = threading.Thread(target=self._timer, daemon=True)
            self._thread.start()

    def _timer(self):
        time.sleep(self.cfg.delay_s)
        self.fire()

    def on_iteration(self, iteration: int):
        """Call once per training iteration (``training.train`` does when ``--inject-fault`` is set)."""
        if self.armed and not self.fired and self.cfg.at_iteration is not None and iteration >= self.cfg.at_iteration:
            self.fire()
        if self.armed and not self.fired and self.cfg.delay_s is not None and self.cfg.start_iteration is not None and iteration >= self.cfg.start_iteration:
            self._start_timer()
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


# ---- function-style API over the reference's flat ``fault_injector_*`` configuration (reference ``fault_injector.py:104-233``) ----
# ``config`` is any object with the ``fault_injector_*`` attributes (the parsed command line, or ``training.config`` objects).
_RNG: Optional[random.Random] = None
_ACTIVE: Optional[FaultInjector] = None


def _rng(config) -> random.Random:
    global _RNG
    if _RNG is None:
        seed = getattr(config, "fault_injector_seed", None)
        _RNG = random.Random(seed if seed is not None else 1234)
    return _RNG


def _csv(v):
    return [x.strip() for x in str(v).split(",") if x.strip()]


def get_fault_ranks(config, world_size: Optional[int] = None):
    """Explicit ``fault_injector_ranks`` ("3" / "1,5") or ``fault_injector_num_ranks`` random ranks — never rank 0, which keeps
    the logs and the rendezvous store alive."""
    import torch.distributed as dist
    world = world_size if world_size is not None else dist.get_world_size()
    forced, n = getattr(config, "fault_injector_ranks", None), getattr(config, "fault_injector_num_ranks", None)
    if forced is not None:
        assert n is None, "give either fault_injector_ranks or fault_injector_num_ranks"
        ranks = [int(r) for r in _csv(forced)]
        assert ranks and all(0 <= r < world for r in ranks), f"fault ranks must lie in [0, {world - 1}]"
        return ranks
    assert n is not None, "give either fault_injector_ranks or fault_injector_num_ranks"
    return _rng(config).sample(range(1, world), k=n)


def get_fault(config) -> Fault:
    kinds = [Fault(k.lower()) if k.lower() in Fault._value2member_map_ else Fault[k.upper()] for k in _csv(getattr(config, "fault_injector_fault_types", None) or "")]
    assert kinds, "fault_injector_fault_types must name at least one fault"
    p = getattr(config, "fault_injector_fault_probabilities", None)
    probs = [float(x) for x in _csv(p)] if p is not None else [1.0] * len(kinds)
    assert len(probs) == len(kinds), "one probability per fault type"
    return _rng(config).choices(kinds, weights=probs, k=1)[0]


def should_setup_fault_injection_at_start(config) -> bool:
    return getattr(config, "fault_injector_delay_start_iteration", None) is None


def should_setup_fault_injection_at_iteration(config, iteration: int) -> bool:
    it = getattr(config, "fault_injector_delay_start_iteration", None)
    return it is not None and it == iteration


def get_fault_delay(config) -> float:
    """Fixed delay, or an exponential inter-arrival draw with mean ``fault_injector_mtti_seconds`` (+ offset): a fleet with a
    given mean time to interrupt."""
    d = getattr(config, "fault_injector_fault_delay", None)
    if d is not None:
        return float(d)
    mtti = getattr(config, "fault_injector_mtti_seconds", None)
    assert mtti is not None, "fault_injector_fault_delay or fault_injector_mtti_seconds must be given"
    import math
    return float(getattr(config, "fault_injector_offset_seconds", None) or 0.0) - math.log(1.0 - _rng(config).random()) * mtti


def setup_fault_injection(config) -> Optional[FaultInjector]:
    """Rank 0 draws the plan (which ranks, which fault, when) and broadcasts it — one tensor: slot r = fault id or NaN, last
    slot = delay — so that the ranks agree even when their RNG streams do not.  Target ranks arm a ``FaultInjector`` (timer
    thread); the training loop polls it (``maybe_raise``) for the exception-type faults."""
    global _ACTIVE
    import math

    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    order = list(Fault)
    dev = torch.device("cuda", torch.cuda.current_device()) if (dist.is_initialized() and dist.get_backend() == "nccl") else torch.device("cpu")
    plan = torch.full((world + 1,), float("nan"), dtype=torch.float64, device=dev)
    if rank == 0:
        fault = get_fault(config)
        for r in get_fault_ranks(config, world):
            plan[r] = float(order.index(fault))
        plan[world] = get_fault_delay(config)
    if world > 1:
        dist.broadcast(plan, src=0)
    mine = float(plan[rank])
    if math.isnan(mine):
        return None
    fault, delay = order[int(mine)], float(plan[world])
    import logging
    logging.getLogger(__name__).warning("FAULT INJECTION: rank %d will inject %s in %.1f s", rank, fault.name, delay)
    _ACTIVE = FaultInjector(FaultInjectorConfig(fault_type=fault, ranks=[rank], delay_s=delay), rank=rank, world_size=world)
    return _ACTIVE


def get_active_fault_injector() -> Optional[FaultInjector]:
    return _ACTIVE
