"""Asynchronous checkpoint writing (reference ``strategies/async_utils.py:237-612``).

A save is split into (1) synchronous staging of device tensors to host memory, (2) the write
itself on a background thread, (3) finalize callbacks (metadata, tracker file) that run on the
main thread once EVERY rank's write has finished."""
from __future__ import annotations

import threading
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
                req._thread.join()
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
