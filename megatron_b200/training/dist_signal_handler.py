"""Cluster-wide view of a POSIX signal (reference ``training/dist_signal_handler.py``): every rank installs the handler, and
``signals_received()`` all-gathers the local flags so all ranks take the same save-and-exit decision in the same iteration."""
from __future__ import annotations

import signal
from typing import List

import torch
import torch.distributed as dist


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu")


class DistributedSignalHandler:
    def __init__(self, sig: int = signal.SIGTERM):
        self.sig = sig
        self._received = False
        self.released = True
        self._original = None

    def signals_received(self) -> List[bool]:
        if not dist.is_initialized():
            return [self._received]
        t = torch.tensor([int(self._received)], device=_device(), dtype=torch.int32)
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [bool(x.item()) for x in out]

    def any_received(self) -> bool:
        return any(self.signals_received())

    def __enter__(self):
        self._received, self.released = False, False
        self._original = signal.getsignal(self.sig)

        def handler(signum, frame):
            self._received = True

        signal.signal(self.sig, handler)
        return self

    def __exit__(self, *exc):
        self.release()

    def release(self) -> bool:
        if self.released:
            return False
        signal.signal(self.sig, self._original)
        self.released = True
        return True
