"""Phase profiler for the RL loop (reference ``megatron/rl/rl_profiling.py``, flags ``--rl-profile`` / ``--rl-profile-dir``).

An RL iteration alternates phases with very different bottlenecks (rollout generation = decode-bound serving; reference / old-policy log-probs = forward-only;
policy update = training), so one tokens/s number says little.  ``RLProfiler.phase(name)`` times a phase on the device (CUDA events on the current stream —
no synchronisation inside the phase — or ``perf_counter`` on CPU), accumulates per-iteration records with user counters (tokens generated, sequences, …), and
``dump`` writes a JSON-lines file plus a per-phase summary (share of the iteration, tokens/s)."""
from __future__ import annotations

import contextlib
import json
import os
import time
from typing import Dict, List, Optional

import torch


class RLProfiler:
    def __init__(self, enabled: bool = True, out_dir: Optional[str] = None, rank: int = 0):
        self.enabled, self.out_dir, self.rank = enabled, out_dir, rank
        self.records: List[Dict] = []
        self._cur: Dict[str, dict] = {}
        self._pending = []            # (name, start event, end event) not yet resolved
        self._iter = 0

    @contextlib.contextmanager
    def phase(self, name: str, **counters):
        if not self.enabled:
            yield
            return
        cuda = torch.cuda.is_available()
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            rec = self._cur.setdefault(name, {"ms": 0.0, "calls": 0})
            rec["calls"] += 1
            for k, v in counters.items():
                rec[k] = rec.get(k, 0) + v
            if cuda:
                e1.record()
                self._pending.append((name, e0, e1))
            else:
                rec["ms"] += (time.perf_counter() - t0) * 1e3

    def count(self, phase: str, **counters) -> None:
        rec = self._cur.setdefault(phase, {"ms": 0.0, "calls": 0})
        for k, v in counters.items():
            rec[k] = rec.get(k, 0) + v

    def end_iteration(self, **extra) -> Dict:
        """Resolve the device timers (one synchronisation per iteration) and close the record."""
        if not self.enabled:
            return {}
        if self._pending:
            torch.cuda.synchronize()
            for name, e0, e1 in self._pending:
                self._cur[name]["ms"] += e0.elapsed_time(e1)
            self._pending.clear()
        total = sum(r["ms"] for r in self._cur.values()) or 1.0
        rec = {"iteration": self._iter, "total_ms": round(total, 3), **extra,
               "phases": {k: {**{kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()}, "share": round(v["ms"] / total, 4)} for k, v in self._cur.items()}}
        self.records.append(rec)
        self._cur = {}
        self._iter += 1
        return rec

    def summary(self) -> Dict[str, dict]:
        out: Dict[str, dict] = {}
        for r in self.records:
            for k, v in r["phases"].items():
                o = out.setdefault(k, {"ms": 0.0, "calls": 0, "tokens": 0})
                o["ms"] += v["ms"]
                o["calls"] += v["calls"]
                o["tokens"] += v.get("tokens", 0)
        total = sum(o["ms"] for o in out.values()) or 1.0
        for o in out.values():
            o["share"] = round(o["ms"] / total, 4)
            o["tokens_per_s"] = round(o["tokens"] / (o["ms"] / 1e3), 1) if o["tokens"] and o["ms"] > 0 else None
            o["ms"] = round(o["ms"], 3)
        return out

    def dump(self, path: Optional[str] = None) -> Optional[str]:
        if not self.enabled or (path is None and self.out_dir is None):
            return None
        path = path or os.path.join(self.out_dir, f"rl_profile_rank{self.rank}.jsonl")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            for r in self.records:
                f.write(json.dumps(r) + "\n")
            f.write(json.dumps({"summary": self.summary()}) + "\n")
        return path
