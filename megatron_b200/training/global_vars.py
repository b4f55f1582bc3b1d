"""Process-wide singletons of the training program (reference ``training/global_vars.py``): args, timers, metric writers, signal handler,
energy monitor.  ``set_global_variables(args)`` builds what the args ask for; getters assert initialisation like the reference's."""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, Optional

import torch.distributed as dist

from ..core.timers import Timers

_G: Dict[str, Any] = {}


def _get(name: str):
    if name not in _G or _G[name] is None:
        raise AssertionError(f"{name} is not initialized")
    return _G[name]


def get_args():
    return _get("args")


def set_args(args) -> None:
    _G["args"] = args


def get_timers() -> Timers:
    return _get("timers")


def get_tensorboard_writer():
    return _G.get("tensorboard")


def get_wandb_writer():
    return _G.get("wandb")


def get_one_logger():
    return _G.get("one_logger")


def get_signal_handler():
    return _get("signal_handler")


def get_energy_monitor():
    return _G.get("energy_monitor")


def get_tokenizer():
    return _get("tokenizer")


class ScalarFileWriter:
    """TensorBoard-compatible ``add_scalar`` / ``add_text`` surface writing JSON lines — used when the ``tensorboard`` package is not
    installed, and by the functional tests, which compare scalars against golden values."""

    def __init__(self, log_dir: str, max_queue: int = 1000, filename: str = "scalars.jsonl"):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, filename)
        self._buf, self._max = [], max_queue
        self._t0 = time.time()

    def add_scalar(self, tag: str, value, global_step: Optional[int] = None, walltime=None):
        self._buf.append({"tag": tag, "value": float(value), "step": global_step, "t": round(time.time() - self._t0, 3)})
        if len(self._buf) >= self._max:
            self.flush()

    def add_text(self, tag: str, text: str, global_step: Optional[int] = None):
        self._buf.append({"tag": tag, "text": text, "step": global_step})

    def log(self, metrics: Dict[str, float], step: Optional[int] = None):       # the W&B surface
        for k, v in metrics.items():
            self.add_scalar(k, v, step)

    def flush(self):
        if self._buf:
            with open(self.path, "a") as f:
                for r in self._buf:
                    f.write(json.dumps(r) + "\n")
            self._buf = []

    def close(self):
        self.flush()

    def read(self):
        self.flush()
        if not os.path.exists(self.path):
            return []
        with open(self.path) as f:
            return [json.loads(l) for l in f]


def _is_last_rank() -> bool:
    return not dist.is_initialized() or dist.get_rank() == dist.get_world_size() - 1


def _build_tensorboard(args):
    d = getattr(args, "tensorboard_dir", None)
    if not d or not _is_last_rank():
        return None
    try:
        from torch.utils.tensorboard import SummaryWriter

        return SummaryWriter(log_dir=d, max_queue=getattr(args, "tensorboard_queue_size", 1000))
    except Exception:
        return ScalarFileWriter(d, getattr(args, "tensorboard_queue_size", 1000))


def _build_wandb(args):
    project = getattr(args, "wandb_project", None)
    if not project or not _is_last_rank():
        return None
    save_dir = getattr(args, "wandb_save_dir", None) or os.path.join(getattr(args, "save", None) or ".", "wandb")
    try:
        import wandb

        wandb.init(dir=save_dir, name=getattr(args, "wandb_exp_name", None), project=project, entity=getattr(args, "wandb_entity", None), config=vars(args), mode=os.environ.get("WANDB_MODE", "offline"))
        return wandb
    except Exception:
        return ScalarFileWriter(save_dir, filename="wandb_offline.jsonl")


def set_global_variables(args, build_tokenizer: bool = False) -> None:
    from .dist_signal_handler import DistributedSignalHandler
    from .one_logger_utils import OneLogger

    _G["args"] = args
    _G["timers"] = Timers(getattr(args, "timing_log_level", 0), getattr(args, "timing_log_option", "minmax"))
    _G["tensorboard"] = _build_tensorboard(args)
    _G["wandb"] = _build_wandb(args)
    _G["one_logger"] = OneLogger(args) if getattr(args, "enable_one_logger", False) else None
    if getattr(args, "exit_signal_handler", False):
        _G["signal_handler"] = DistributedSignalHandler().__enter__()
    if getattr(args, "log_energy", False):
        from .energy_monitor import EnergyMonitor

        _G["energy_monitor"] = EnergyMonitor()
    if build_tokenizer:
        from ..core.tokenizers import build_tokenizer as bt

        _G["tokenizer"] = bt(args)


def unset_global_variables() -> None:
    for w in ("tensorboard", "wandb"):
        if _G.get(w) is not None and hasattr(_G[w], "close"):
            try:
                _G[w].close()
            except Exception:
                pass
    _G.clear()


def destroy_global_vars() -> None:
    unset_global_variables()
