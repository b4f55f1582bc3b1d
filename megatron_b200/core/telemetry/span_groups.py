"""Span tracing with a no-op fallback (reference ``core/telemetry/{span_groups,fallbacks}.py``: OpenTelemetry spans for job / startup / train / iteration /
forward-backward / optimizer, per-layer and p2p spans).

``span("train.iteration", iteration=i)`` is a context manager.  Sinks, in order of preference: an OpenTelemetry tracer when the ``opentelemetry`` API is
importable and a provider is configured; otherwise the in-process ``SpanRecorder`` (bounded ring of finished spans + JSON-lines export), which is also what the
tests use; when tracing is disabled the context manager is a shared no-op object, so instrumented hot paths cost one attribute check."""
from __future__ import annotations

import json
import os
import threading
import time
from collections import deque
from contextlib import contextmanager
from typing import Deque, Dict, List, Optional

SPAN_GROUPS = {
    "job": ("job",),
    "startup": ("startup.initialize", "startup.model_setup", "startup.data_setup", "startup.load_checkpoint"),
    "train": ("train.loop", "train.iteration", "train.forward_backward", "train.optimizer_step", "train.eval", "train.save_checkpoint"),
    "layer": ("layer.attention", "layer.mlp"),
    "comm": ("p2p.send_recv", "collective"),
}


class SpanRecorder:
    def __init__(self, capacity: int = 4096, path: Optional[str] = None, enabled_groups: Optional[List[str]] = None):
        self.finished: Deque[dict] = deque(maxlen=capacity)
        self.path = path
        self.enabled = set(enabled_groups) if enabled_groups is not None else {"job", "startup", "train"}
        self._stack = threading.local()
        self._lock = threading.Lock()

    def group_enabled(self, name: str) -> bool:
        head = name.split(".", 1)[0]
        return head in self.enabled or any(name in v for k, v in SPAN_GROUPS.items() if k in self.enabled)

    @contextmanager
    def span(self, name: str, **attrs):
        if not self.group_enabled(name):
            yield None
            return
        st = getattr(self._stack, "s", None)
        if st is None:
            st = self._stack.s = []
        rec = {"name": name, "start": time.time(), "parent": st[-1]["name"] if st else None, "attrs": attrs}
        st.append(rec)
        try:
            yield rec
        except BaseException as e:
            rec["error"] = type(e).__name__
            raise
        finally:
            st.pop()
            rec["duration_s"] = time.time() - rec["start"]
            with self._lock:
                self.finished.append(rec)
                if self.path:
                    with open(self.path, "a") as f:
                        f.write(json.dumps(rec, default=str) + "\n")

    def summary(self) -> Dict[str, dict]:
        out: Dict[str, dict] = {}
        for r in self.finished:
            d = out.setdefault(r["name"], {"count": 0, "total_s": 0.0, "max_s": 0.0})
            d["count"] += 1
            d["total_s"] += r["duration_s"]
            d["max_s"] = max(d["max_s"], r["duration_s"])
        return out


class _Noop:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOOP = _Noop()
_RECORDER: Optional[SpanRecorder] = None
_OTEL = None


def set_recorder(rec: Optional[SpanRecorder]) -> None:
    global _RECORDER
    _RECORDER = rec


def get_recorder() -> Optional[SpanRecorder]:
    return _RECORDER


def enable_from_env() -> Optional[SpanRecorder]:
    """``MEGATRON_B200_TRACE=<file.jsonl>[:group,group]`` turns the recorder on (OpenTelemetry is used instead when a tracer provider is configured)."""
    global _OTEL
    spec = os.environ.get("MEGATRON_B200_TRACE")
    if not spec:
        return None
    path, _, groups = spec.partition(":")
    try:
        from opentelemetry import trace

        _OTEL = trace.get_tracer("megatron_b200")
    except Exception:
        _OTEL = None
    set_recorder(SpanRecorder(path=path or None, enabled_groups=groups.split(",") if groups else None))
    return _RECORDER


def span(name: str, **attrs):
    if _RECORDER is None:
        return _NOOP
    if _OTEL is not None and _RECORDER.group_enabled(name):
        return _OTEL.start_as_current_span(name, attributes={k: v for k, v in attrs.items() if isinstance(v, (str, int, float, bool))})
    return _RECORDER.span(name, **attrs)
