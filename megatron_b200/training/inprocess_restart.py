"""In-process restart + fault-tolerance hooks (reference ``training/inprocess_restart.py`` and ``ft_integration.py``, which wrap
nvidia-resiliency-ext).  Self-contained equivalent:

``maybe_wrap_for_inprocess_restart(pretrain_fn)`` returns a function that, on a recoverable failure (exception in the training thread —
including ``InjectedFaultError`` from ``core.fault_injector`` and NaN-triggered reruns), tears the model-parallel state down, re-creates
the process groups and calls ``pretrain_fn`` again; training resumes from the last checkpoint (``--load``).  Bounded by
``--inprocess-max-iterations``.  Heartbeat timeouts are a thread that aborts the process (so the launcher restarts it) when the training
loop has not called ``heartbeat()`` for ``timeout_s``."""
from __future__ import annotations

import os
import sys
import threading
import time
import traceback
from typing import Callable, Optional

import torch


class Heartbeat:
    def __init__(self, timeout_s: float, on_timeout: Optional[Callable[[], None]] = None):
        self.timeout_s, self._last, self._stop = timeout_s, time.time(), False
        self._on_timeout = on_timeout or (lambda: os._exit(66))
        self._thread = threading.Thread(target=self._watch, daemon=True)

    def start(self):
        self._thread.start()
        return self

    def beat(self):
        self._last = time.time()

    def stop(self):
        self._stop = True

    def _watch(self):
        while not self._stop:
            time.sleep(min(1.0, self.timeout_s / 4))
            if time.time() - self._last > self.timeout_s:
                print(f"[ft] no heartbeat for {self.timeout_s}s — aborting so the launcher can restart this rank", file=sys.stderr, flush=True)
                self._on_timeout()
                return


_HEARTBEAT: Optional[Heartbeat] = None


def setup(timeout_s: Optional[float] = None):
    """``ft_integration.setup`` equivalent; enabled with ``--enable-ft-package`` / ``timeout_s``."""
    global _HEARTBEAT
    if timeout_s:
        _HEARTBEAT = Heartbeat(timeout_s).start()


def heartbeat():
    if _HEARTBEAT is not None:
        _HEARTBEAT.beat()


def shutdown():
    if _HEARTBEAT is not None:
        _HEARTBEAT.stop()


def _teardown():
    try:
        from .training import destroy_global_state

        destroy_global_state()          # args / timers / writers, micro-batch calculator, rerun state machine, model-parallel groups, async-save queue
    except Exception:
        from ..core import parallel_state as ps

        try:
            ps.destroy_model_parallel()
        except Exception:
            pass
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def maybe_wrap_for_inprocess_restart(pretrain_fn: Callable, max_restarts: Optional[int] = None, recoverable=(RuntimeError, FloatingPointError)):
    max_restarts = int(os.environ.get("MEGATRON_B200_INPROCESS_RESTARTS", "0")) if max_restarts is None else max_restarts
    if max_restarts <= 0:
        return pretrain_fn

    def wrapped(*args, **kwargs):
        attempt = 0
        while True:
            try:
                return pretrain_fn(*args, **kwargs)
            except recoverable as e:
                attempt += 1
                print(f"[inprocess-restart] attempt {attempt}/{max_restarts} after {type(e).__name__}: {e}\\n{traceback.format_exc(limit=3)}", file=sys.stderr, flush=True)
                if attempt > max_restarts:
                    raise
                _teardown()
                if torch.distributed.is_initialized():
                    torch.distributed.barrier()

    return wrapped
