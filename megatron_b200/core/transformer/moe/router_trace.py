"""Routing-decision tracing to disk (reference ``transformer/moe/router_trace.py:45-492``).

Forward hooks on every router write, per step and per layer, the top-k expert ids of each
token — optionally with the router input and the logits — so routing collapse, drift
between training and serving, or expert hot-spots can be analysed offline (and fed back
through ``RouterReplay``).  Records are buffered on the host and flushed as one
``index.jsonl`` line + one ``.pt`` payload per record; only ranks selected by
``trace_ranks`` write."""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional, Tuple

import torch

_TRACER: Optional["RouterTracer"] = None


def _parse_router_module_name(module_name: str) -> Optional[Tuple[str, Optional[int], int]]:
    """``decoder.layers.3.mlp.router`` → ``("decoder", None, 3)``; MTP routers carry their depth."""
    m = re.search(r"mtp\.layers\.(\d+)\..*?(?:layers?\.(\d+)\.)?mlp\.router$", module_name)
    if m:
        return "mtp", int(m.group(1)), int(m.group(2) or 0)
    m = re.search(r"(\w+)\.layers\.(\d+)\.mlp\.router$", module_name)
    if m:
        return m.group(1), None, int(m.group(2))
    return None


def init_moe_router_tracer(trace_dir: str, save_hidden_states: bool = False, save_logits: bool = False, flush_every: int = 64,
                           trace_ranks: Optional[List[int]] = None, start_step: int = 0, end_step: Optional[int] = None) -> "RouterTracer":
    global _TRACER
    _TRACER = RouterTracer(trace_dir, save_hidden_states, save_logits, flush_every, trace_ranks, start_step, end_step)
    return _TRACER


def get_moe_router_tracer() -> Optional["RouterTracer"]:
    return _TRACER


def _load(record: dict, trace_dir: str, field: str) -> torch.Tensor:
    if field not in record.get("fields", []):
        raise KeyError(f"record has no {field!r} (tracer was created without it)")
    return torch.load(os.path.join(trace_dir, record["file"]), map_location="cpu", weights_only=True)[field]


def load_hidden_states_for_record(record: dict, trace_dir: str) -> torch.Tensor:
    return _load(record, trace_dir, "hidden_states")


def load_logits_for_record(record: dict, trace_dir: str) -> torch.Tensor:
    return _load(record, trace_dir, "logits")


def load_indices_for_record(record: dict, trace_dir: str) -> torch.Tensor:
    return _load(record, trace_dir, "top_indices")


class RouterTracer:
    def __init__(self, trace_dir: str, save_hidden_states: bool = False, save_logits: bool = False, flush_every: int = 64,
                 trace_ranks: Optional[List[int]] = None, start_step: int = 0, end_step: Optional[int] = None):
        import torch.distributed as dist

        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.enabled = trace_ranks is None or self.rank in trace_ranks
        self.trace_dir = os.path.join(trace_dir, f"rank_{self.rank:05d}")
        self.save_hidden_states, self.save_logits = save_hidden_states, save_logits
        self.flush_every = max(1, flush_every)
        self.start_step, self.end_step = start_step, end_step
        self.step = 0
        self._calls: Dict[str, int] = {}           # microbatch counter per router within the step
        self._pending: List[Tuple[dict, dict]] = []
        self._handles = []
        if self.enabled:
            os.makedirs(self.trace_dir, exist_ok=True)

    # ---- hooks ----------------------------------------------------------------------------
    def register_hooks(self, model) -> None:
        from .router import Router

        models = model if isinstance(model, (list, tuple)) else [model]
        for chunk, m in enumerate(models):
            for name, mod in m.named_modules():
                if isinstance(mod, Router):
                    self._handles.append(mod.register_forward_hook(self.make_hook(f"chunk{chunk}.{name}" if len(models) > 1 else name)))

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()

    def advance_step(self, step_id: Optional[int] = None) -> None:
        self.step = self.step + 1 if step_id is None else step_id
        self._calls.clear()
        if len(self._pending) >= self.flush_every:
            self._flush_records_to_disk()

    def _active(self) -> bool:
        return self.enabled and self.step >= self.start_step and (self.end_step is None or self.step < self.end_step)

    def make_hook(self, module_name: str = ""):
        def hook(module, inputs, outputs):
            if self._active():
                self._record(module, inputs, outputs, module_name)
        return hook

    @staticmethod
    def _extract_hidden_state(inputs, expected_num_tokens):
        x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
        if not torch.is_tensor(x):
            return None
        x = x.reshape(-1, x.shape[-1])
        return x if x.shape[0] == expected_num_tokens else None

    def _record(self, module, inputs, outputs, module_name="") -> None:
        probs, routing_map = outputs[0], outputs[1]
        topk = int(getattr(module, "topk", routing_map.sum(-1).max().item()))
        # dense [T, E] map → [T, k] ids, highest probability first (dropped slots → -1)
        masked = torch.where(routing_map, probs.detach().float(), torch.full_like(probs, float("-inf"), dtype=torch.float32))
        vals, idx = masked.topk(topk, dim=-1)
        idx = torch.where(torch.isfinite(vals), idx, torch.full_like(idx, -1))
        payload = {}
        if self.save_hidden_states:
            h = self._extract_hidden_state(inputs, idx.shape[0])
            if h is not None:
                payload["hidden_states"] = h.detach().to("cpu")
        if self.save_logits:
            x = inputs[0].reshape(-1, inputs[0].shape[-1])
            with torch.no_grad():
                payload["logits"] = module.gating(x).detach().float().to("cpu")
        self.record_indices(idx, module_name=module_name, layer_number=getattr(module, "layer_number", None), extra=payload)

    def record_indices(self, top_indices: torch.Tensor, module_name: str = "", layer_number: Optional[int] = None, extra: Optional[dict] = None) -> None:
        """Public entry for callers that already hold ``[T, k]`` ids (fused routers, inference)."""
        if not self._active():
            return
        parsed = _parse_router_module_name(module_name) or ("", None, (layer_number or 0) - 1)
        mb = self._calls.get(module_name, 0)
        self._calls[module_name] = mb + 1
        payload = dict(extra or {})
        payload["top_indices"] = top_indices.detach().to("cpu", torch.int32)
        rec = {"step": self.step, "microbatch": mb, "module": module_name, "block": parsed[0], "mtp_index": parsed[1], "layer": parsed[2],
               "layer_number": layer_number, "num_tokens": int(top_indices.shape[0]), "topk": int(top_indices.shape[1]), "fields": sorted(payload),
               "file": f"s{self.step:08d}_{len(self._pending):06d}_{abs(hash(module_name)) % 10**8:08d}_mb{mb}.pt"}
        self._pending.append((rec, payload))
        if len(self._pending) >= self.flush_every:
            self._flush_records_to_disk()

    # ---- storage --------------------------------------------------------------------------
    def _flush_records_to_disk(self) -> None:
        if not self._pending:
            return
        with open(os.path.join(self.trace_dir, "index.jsonl"), "a") as f:
            for rec, payload in self._pending:
                torch.save(payload, os.path.join(self.trace_dir, rec["file"]))
                f.write(json.dumps(rec) + "\n")
        self._pending.clear()

    def flush(self) -> None:
        self._flush_records_to_disk()

    def read_index(self) -> List[dict]:
        self.flush()
        p = os.path.join(self.trace_dir, "index.jsonl")
        return [json.loads(l) for l in open(p)] if os.path.exists(p) else []
