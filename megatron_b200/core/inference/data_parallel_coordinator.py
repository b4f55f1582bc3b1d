"""Data-parallel inference coordinator (reference ``inference/data_parallel_inference_coordinator.py``: a ZMQ router in front of one dynamic engine per
DP replica).  Requests are routed to the replica with the least outstanding work (prompt + expected generation tokens), finished requests flow back
tagged with a global id, and paused / drained replicas are skipped.  The transport is pluggable: in-process engines (tests, single-node serving
where all replicas live in one launcher) or any object with ``add_request`` / ``step`` / ``has_unfinished`` (e.g. a proxy over a
``multiprocessing`` pipe or an HTTP client to ``tools/run_text_generation_server.py``)."""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .sampling import SamplingParams


@dataclass
class _Replica:
    engine: object
    outstanding_tokens: int = 0
    local_to_global: Dict[int, int] = field(default_factory=dict)
    cost: Dict[int, int] = field(default_factory=dict)
    paused: bool = False
    served: int = 0


class DataParallelInferenceCoordinator:
    def __init__(self, engines: List[object]):
        assert engines, "need at least one engine"
        self.replicas = [_Replica(e) for e in engines]
        self._ids = itertools.count()
        self.finished: Dict[int, object] = {}
        self.placement: Dict[int, int] = {}

    def pause(self, replica: int) -> None:
        self.replicas[replica].paused = True

    def resume(self, replica: int) -> None:
        self.replicas[replica].paused = False

    def add_request(self, prompt_tokens: List[int], sampling_params: Optional[SamplingParams] = None) -> int:
        sp = sampling_params or SamplingParams()
        cands = [(r.outstanding_tokens, i) for i, r in enumerate(self.replicas) if not r.paused]
        if not cands:
            raise RuntimeError("all data-parallel replicas are paused")
        _, i = min(cands)
        rep = self.replicas[i]
        gid = next(self._ids)
        lid = rep.engine.add_request(prompt_tokens, sp)
        cost = len(prompt_tokens) + sp.num_tokens_to_generate
        rep.local_to_global[lid], rep.cost[lid] = gid, cost
        rep.outstanding_tokens += cost
        self.placement[gid] = i
        return gid

    def has_unfinished(self) -> bool:
        return any(r.engine.has_unfinished() for r in self.replicas)

    def step(self) -> List[int]:
        """One engine step on every replica that has work; returns the global ids that finished in this step."""
        done = []
        for rep in self.replicas:
            if not rep.engine.has_unfinished():
                continue
            for req in rep.engine.step():
                gid = rep.local_to_global.pop(req.request_id)
                rep.outstanding_tokens -= rep.cost.pop(req.request_id)
                rep.served += 1
                self.finished[gid] = req
                done.append(gid)
        return done

    def run_until_done(self) -> Dict[int, object]:
        while self.has_unfinished():
            self.step()
        return self.finished
