"""Asynchronous checkpoint writing (reference ``strategies/async_utils.py:237-612``, ``filesystem_async.py:114-440``).

A save is split into (1) synchronous, collective planning + staging of the shards to host memory, (2) the write itself in the background — on a thread
(``AsyncRequest``) or in a persistent worker PROCESS (``ProcessAsyncRequest`` / ``PersistentWriterProcess``: no GIL contention with the training loop, the
shards travel as shared-memory tensors) —, (3) finalize callbacks (metadata commit, tracker file) on the main thread once EVERY rank's write has finished."""
from __future__ import annotations

import multiprocessing as _mp
import threading
import traceback
from collections import deque
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


class AsyncRequest:
    def __init__(self, async_fn: Optional[Callable], async_fn_args: Tuple, finalize_fns: List[Callable]):
        self.async_fn, self.async_fn_args, self.finalize_fns = async_fn, async_fn_args, list(finalize_fns)
        self.is_frozen = False
        self._thread: Optional[threading.Thread] = None
        self._exc: Optional[BaseException] = None

    def add_finalize_fn(self, fn: Callable):
        if self.is_frozen:
            raise RuntimeError("cannot add finalize functions to a frozen AsyncRequest")
        self.finalize_fns.append(fn)

    def freeze(self) -> "AsyncRequest":
        self.is_frozen = True
        return self

    def start(self):
        def run():
            try:
                if self.async_fn is not None:
                    self.async_fn(*self.async_fn_args)
            except BaseException as e:  # noqa: BLE001
                self._exc = e

        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()

    def is_done(self) -> bool:
        return self._thread is not None and not self._thread.is_alive()

    def execute_sync(self):
        if self.async_fn is not None:
            self.async_fn(*self.async_fn_args)
        self.finalize()

    def finalize(self):
        if self._thread is not None:
            self._thread.join()
        if self._exc is not None:
            raise self._exc
        for fn in self.finalize_fns:
            fn()


class AsyncCallsQueue:
    """FIFO of in-flight saves; ``maybe_finalize_async_calls`` is polled from the training loop."""

    def __init__(self, persistent: bool = False):
        self.q = deque()
        self.idx = 0

    def schedule_async_request(self, req: AsyncRequest) -> int:
        req.freeze()
        req.start()
        self.idx += 1
        self.q.append((self.idx, req))
        return self.idx

    def _all_ranks_done(self, req: AsyncRequest) -> bool:
        done = req.is_done()
        if dist.is_available() and dist.is_initialized():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([1 if done else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item())
        return done

    def maybe_finalize_async_calls(self, blocking: bool = False, no_dist: bool = False) -> List[int]:
        finished = []
        while self.q:
            idx, req = self.q[0]
            if blocking:
                if req._thread is not None:
                    req._thread.join()
                elif hasattr(req, "_poll"):
                    req._poll(True)
            if not (req.is_done() if no_dist else self._all_ranks_done(req)):
                break
            req.finalize()
            if dist.is_available() and dist.is_initialized() and not no_dist:
                dist.barrier()
            self.q.popleft()
            finished.append(idx)
        return finished

    def get_num_unfinalized_calls(self) -> int:
        return len(self.q)

    def close(self):
        self.maybe_finalize_async_calls(blocking=True)



# ---- worker-process variant ---------------------------------------------------------------------------------------------------------------------------------
def _writer_main(jobs, results):
    """Loop of the persistent writer process: (job id, fn, args) -> (job id, ok, value | traceback)."""
    torch.set_num_threads(1)
    while True:
        job = jobs.get()
        if job is None:
            return
        jid, fn, args = job
        try:
            results.put((jid, True, fn(*args)))
        except BaseException:  # noqa: BLE001
            results.put((jid, False, traceback.format_exc()))


class PersistentWriterProcess:
    """One spawned worker per training process, reused by every checkpoint (reference ``PersistentAsyncCaller``).  Spawn — not fork — so the child never
    inherits CUDA / NCCL state."""

    _instance: Optional["PersistentWriterProcess"] = None

    def __init__(self):
        import torch.multiprocessing as tmp

        ctx = tmp.get_context("spawn")
        self.jobs, self.results = ctx.Queue(), ctx.Queue()
        self.proc = ctx.Process(target=_writer_main, args=(self.jobs, self.results), daemon=True)
        self.proc.start()
        self._next = 0
        self._done = {}

    @classmethod
    def get(cls) -> "PersistentWriterProcess":
        if cls._instance is None or not cls._instance.proc.is_alive():
            cls._instance = cls()
        return cls._instance

    def submit(self, fn: Callable, args: Tuple) -> int:
        self._next += 1
        self.jobs.put((self._next, fn, args))
        return self._next

    def poll(self, jid: int, block: bool = False):
        """None while running; (ok, value) when finished."""
        while jid not in self._done:
            try:
                j, ok, val = self.results.get(block, 600 if block else None) if block else self.results.get_nowait()
            except Exception:
                if not self.proc.is_alive():
                    return (False, "checkpoint writer process died")
                if not block:
                    return None
                continue
            self._done[j] = (ok, val)
        return self._done.pop(jid)

    def close(self):
        try:
            self.jobs.put(None)
            self.proc.join(timeout=10)
        except Exception:
            pass
        PersistentWriterProcess._instance = None


class ProcessAsyncRequest(AsyncRequest):
    """``async_fn(*args)`` runs in the persistent writer process; its return value is handed to ``finalize_fns`` that accept one argument."""

    def __init__(self, async_fn: Callable, async_fn_args: Tuple, finalize_fns: List[Callable]):
        super().__init__(async_fn, async_fn_args, finalize_fns)
        self._jid: Optional[int] = None
        self._result = None
        self._have_result = False

    def start(self):
        self._jid = PersistentWriterProcess.get().submit(self.async_fn, self.async_fn_args)

    def _poll(self, block: bool):
        if self._have_result or self._jid is None:
            return
        r = PersistentWriterProcess.get().poll(self._jid, block)
        if r is not None:
            ok, val = r
            self._have_result = True
            if ok:
                self._result = val
            else:
                self._exc = RuntimeError(f"asynchronous checkpoint write failed in the writer process:\n{val}")

    def is_done(self) -> bool:
        self._poll(False)
        return self._have_result

    def execute_sync(self):
        self._result = self.async_fn(*self.async_fn_args)
        self._have_result = True
        self.finalize()

    def finalize(self):
        self._poll(True)
        if self._exc is not None:
            raise self._exc
        import inspect

        for fn in self.finalize_fns:
            fn(self._result) if len(inspect.signature(fn).parameters) >= 1 else fn()
