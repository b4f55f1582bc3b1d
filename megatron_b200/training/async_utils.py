"""Async-checkpoint plumbing for the training loop (reference ``training/async_utils.py``): one process-wide queue of in-flight saves,
finalised either opportunistically every iteration or blockingly at exit."""
from __future__ import annotations

from typing import Optional

from ..core.dist_checkpointing.strategies.async_utils import AsyncCallsQueue, AsyncRequest

_QUEUE: Optional[AsyncCallsQueue] = None


def init_persistent_async_worker() -> AsyncCallsQueue:
    """Create the queue up front (the worker thread/process is started once and reused by every save)."""
    global _QUEUE
    if _QUEUE is None:
        _QUEUE = AsyncCallsQueue()
    return _QUEUE


def schedule_async_save(request: AsyncRequest) -> int:
    return init_persistent_async_worker().schedule_async_request(request)


def maybe_finalize_async_save(blocking: bool = False, terminate: bool = False) -> None:
    global _QUEUE
    if _QUEUE is None:
        return
    _QUEUE.maybe_finalize_async_calls(blocking)
    if terminate:
        _QUEUE.close()
        _QUEUE = None


def is_empty_async_queue() -> bool:
    return _QUEUE is None or _QUEUE.get_num_unfinalized_calls() == 0


def reset_persistent_async_worker() -> None:
    maybe_finalize_async_save(blocking=True, terminate=True)
