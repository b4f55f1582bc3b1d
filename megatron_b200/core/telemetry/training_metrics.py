"""Metric instruments of the training loop (reference ``core/telemetry/training_metrics.py``; names follow ``docs/user-guide/observability/metrics.md``):
counters / gauges / histograms behind a tiny facade that exports to OpenTelemetry when a meter provider is configured, to Prometheus
(``prometheus_client`` is in the image) when ``start_http_server`` is requested, and always keeps the last values in memory for the loop's own logging."""
from __future__ import annotations

from typing import Dict, Optional

METRIC_NAMES = ("train.iteration_time_ms", "train.tokens_per_second", "train.tflops_per_gpu", "train.lm_loss", "train.grad_norm", "train.learning_rate",
                "train.loss_scale", "train.skipped_iterations", "train.consumed_samples", "train.checkpoint_save_seconds", "train.energy_joules_per_iteration",
                "train.memory_allocated_gib")


class TrainingMetrics:
    def __init__(self, prometheus_port: Optional[int] = None, labels: Optional[Dict[str, str]] = None):
        self.values: Dict[str, float] = {}
        self.labels = labels or {}
        self._prom = {}
        self._otel = {}
        if prometheus_port is not None:
            try:
                import prometheus_client as pc

                pc.start_http_server(prometheus_port)
                self._prom = {n: pc.Gauge(n.replace(".", "_"), n, list(self.labels)) for n in METRIC_NAMES}
            except Exception:
                self._prom = {}
        try:
            from opentelemetry import metrics

            meter = metrics.get_meter("megatron_b200")
            self._otel = {n: meter.create_gauge(n) if hasattr(meter, "create_gauge") else meter.create_histogram(n) for n in METRIC_NAMES}
        except Exception:
            self._otel = {}

    def record(self, name: str, value: float) -> None:
        self.values[name] = float(value)
        g = self._prom.get(name)
        if g is not None:
            (g.labels(**self.labels) if self.labels else g).set(float(value))
        o = self._otel.get(name)
        if o is not None:
            try:
                (o.set if hasattr(o, "set") else o.record)(float(value), attributes=self.labels)
            except Exception:
                pass

    def record_iteration(self, *, iteration_time_s: float, tokens: int, flops: float, world: int, loss: Optional[float] = None, grad_norm: Optional[float] = None,
                         lr: Optional[float] = None, loss_scale: Optional[float] = None) -> None:
        self.record("train.iteration_time_ms", iteration_time_s * 1e3)
        self.record("train.tokens_per_second", tokens / max(iteration_time_s, 1e-9))
        self.record("train.tflops_per_gpu", flops / max(iteration_time_s, 1e-9) / 1e12 / max(world, 1))
        for n, v in (("train.lm_loss", loss), ("train.grad_norm", grad_norm), ("train.learning_rate", lr), ("train.loss_scale", loss_scale)):
            if v is not None:
                self.record(n, v)
