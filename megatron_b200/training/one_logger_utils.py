"""End-to-end job metrics (reference ``training/one_logger_utils.py``): app / train-loop / checkpoint / eval timings and throughput that a
cluster-level dashboard consumes.  Here the sink is a JSON-lines file (and the W&B writer when present); the tracked keys follow the reference."""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, Optional


def _ms() -> int:
    return round(time.time() * 1000.0)


class OneLogger:
    def __init__(self, args=None, path: Optional[str] = None):
        self.metrics: Dict[str, Any] = {"app_start_time": _ms()}
        self.path = path or os.path.join(getattr(args, "save", None) or ".", "one_logger.jsonl")
        self._ckpt_t0 = None
        self._args = args

    def store_set(self, key: str, value: Any) -> None:
        self.metrics[key] = value

    def store_get(self, key: str, default=None):
        return self.metrics.get(key, default)

    def store_has_key(self, key: str) -> bool:
        return key in self.metrics

    def log_metrics(self, m: Dict[str, Any]) -> None:
        self.metrics.update(m)
        try:
            os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
            with open(self.path, "a") as f:
                f.write(json.dumps(m) + "\n")
        except OSError:
            pass

    # ---- callbacks used by the training loop ----
    def on_train_start(self, iteration: int, consumed_train_samples: int, train_samples: int, seq_length: int, train_iterations: int, save: Optional[str],
                       async_save: bool, log_throughput: bool, num_floating_point_operations_so_far: float) -> None:
        self.log_metrics({
            "train_iterations_start": iteration, "train_samples_start": consumed_train_samples, "train_iterations_target": train_iterations,
            "train_samples_target": train_samples, "train_tokens_target": seq_length * train_samples, "app_train_loop_start_time": _ms(),
            "is_save_checkpoint_enabled": bool(save), "save_checkpoint_strategy": "async" if async_save else "sync", "is_log_throughput_enabled": log_throughput,
            "train_tflop_start": num_floating_point_operations_so_far / 1e12,
        })
        self.metrics.update(train_iterations_time_msecs_total=0.0, tracked_train_iterations=0, save_checkpoint_count=0, save_checkpoint_sync_time_total=0.0)

    def track_iteration(self, elapsed_s: float, global_batch_size: int, seq_length: int, flops: float) -> None:
        m = self.metrics
        m["train_iterations_time_msecs_total"] = m.get("train_iterations_time_msecs_total", 0.0) + elapsed_s * 1e3
        m["tracked_train_iterations"] = m.get("tracked_train_iterations", 0) + 1
        m["train_samples_end"] = m.get("train_samples_end", m.get("train_samples_start", 0)) + global_batch_size
        m["train_tflop_end"] = m.get("train_tflop_end", m.get("train_tflop_start", 0.0)) + flops / 1e12
        n = m["tracked_train_iterations"]
        m["train_iterations_time_msecs_avg"] = m["train_iterations_time_msecs_total"] / n
        m["train_throughput_per_gpu"] = flops / 1e12 / max(elapsed_s, 1e-9) / max(getattr(self._args, "world_size", 1) or 1, 1)

    def on_save_checkpoint_start(self, async_save: bool) -> None:
        self._ckpt_t0 = time.time()
        self.metrics["save_checkpoint_count"] = self.metrics.get("save_checkpoint_count", 0) + 1

    def on_save_checkpoint_end(self, iteration: int, async_save: bool) -> None:
        dt = time.time() - (self._ckpt_t0 or time.time())
        m = self.metrics
        m["save_checkpoint_sync_time_total"] = m.get("save_checkpoint_sync_time_total", 0.0) + dt
        m["save_checkpoint_sync_time_max"] = max(m.get("save_checkpoint_sync_time_max", 0.0), dt)
        m["save_checkpoint_sync_time_min"] = min(m.get("save_checkpoint_sync_time_min", float("inf")), dt)
        if not async_save:
            self.on_save_checkpoint_success(iteration)

    def on_save_checkpoint_success(self, iteration: int) -> None:
        self.log_metrics({"last_successful_save_checkpoint_iteration": iteration, "last_successful_save_checkpoint_time": _ms()})

    def on_eval(self, iteration: int, elapsed_s: float) -> None:
        self.log_metrics({"validation_iterations_time_msecs": elapsed_s * 1e3, "tracked_validation_iteration": iteration})

    def on_train_end(self) -> None:
        keys = ("train_iterations_time_msecs_total", "train_iterations_time_msecs_avg", "tracked_train_iterations", "train_samples_end", "train_tflop_end",
                "train_throughput_per_gpu", "save_checkpoint_count", "save_checkpoint_sync_time_total")
        self.log_metrics({"app_train_loop_finish_time": _ms(), **{k: self.metrics[k] for k in keys if k in self.metrics}})

    def finish(self) -> None:
        self.log_metrics({"app_finish_time": _ms()})


def on_save_checkpoint_start(async_save: bool) -> None:
    from .global_vars import get_one_logger

    ol = get_one_logger()
    if ol is not None:
        ol.on_save_checkpoint_start(async_save)


def on_save_checkpoint_end(iteration: int, async_save: bool) -> None:
    from .global_vars import get_one_logger

    ol = get_one_logger()
    if ol is not None:
        ol.on_save_checkpoint_end(iteration, async_save)
