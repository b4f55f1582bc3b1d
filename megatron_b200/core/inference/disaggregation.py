"""Prefill / decode disaggregation (reference ``inference/disaggregation/``: separate prefill and decode workers with NCCL / NIXL KV transfer).

Prefill is compute-bound (long GEMMs over the prompt), decode is bandwidth- and latency-bound (one token per step against the whole KV cache); running them
on different GPU pools lets each be batched and parallelised for its own regime and stops long prompts from stalling token streaming.  Pieces:

* ``PrefillWorker``  — wraps a ``DynamicInferenceEngine``: runs ONLY the prompt forward of a request, samples the first token, and exports the request's KV pages.
* ``KVTransport``    — moves ``KVPayload`` (per-layer K/V of the prompt + request metadata) between workers: ``InProcessTransport`` (same process, tests),
  ``TorchDistTransport`` (``torch.distributed`` point-to-point — NCCL on GPUs, i.e. NVLink inside the box; gloo on CPU).  Payloads are sent as ONE flat tensor per
  direction plus a small header, so a transfer is a single large copy.
* ``DecodeWorker``   — admits a transferred request directly into the running set of its engine (pages allocated, KV imported, first token already known) and
  decodes with continuous batching as usual.

The outputs are identical to a single engine's (same kernels, same sampling); only WHERE the two phases run changes."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .engine import DynamicInferenceEngine, InferenceRequest
from .sampling import SamplingParams


@dataclass
class KVPayload:
    request_id: int
    prompt_tokens: List[int]
    first_token: int
    sampling_params: SamplingParams
    k: torch.Tensor          # [layers, prompt_len, kv_heads, head_dim]
    v: torch.Tensor

    @property
    def nbytes(self) -> int:
        return 2 * self.k.numel() * self.k.element_size()


class PrefillWorker:
    def __init__(self, engine: DynamicInferenceEngine):
        self.engine = engine

    @torch.no_grad()
    def prefill(self, request_id: int, prompt_tokens: List[int], sampling_params: Optional[SamplingParams] = None) -> KVPayload:
        eng, sp = self.engine, sampling_params or SamplingParams()
        eng.model.eval()
        req = InferenceRequest(request_id, list(prompt_tokens), sp)
        if not eng.cache.add_request(request_id, len(prompt_tokens)):
            raise MemoryError("prefill worker is out of KV pages")
        try:
            logits = eng._forward_request(req, req.prompt_tokens, 0)
            eng._emit(req, logits)
            n = len(prompt_tokens)
            ks, vs = zip(*(eng.cache.gather(li, request_id, n) for li in range(eng.cache.k.shape[0])))
            return KVPayload(request_id, list(prompt_tokens), req.generated_tokens[0], sp, torch.stack(ks).clone(), torch.stack(vs).clone())
        finally:
            eng.cache.release(request_id)      # the pages live on the decode side from now on


class DecodeWorker:
    def __init__(self, engine: DynamicInferenceEngine):
        self.engine = engine

    def admit(self, payload: KVPayload) -> bool:
        """Import a prefilled request; returns False when there are not enough free pages right now (caller retries after a step)."""
        eng, sp = self.engine, payload.sampling_params
        n = len(payload.prompt_tokens)
        if not eng.cache.can_admit(n + sp.num_tokens_to_generate) or not eng.cache.add_request(payload.request_id, n):
            return False
        for li in range(payload.k.shape[0]):
            eng.cache.append(li, payload.request_id, payload.k[li].to(eng.device), payload.v[li].to(eng.device), 0)
        eng.cache.lengths[payload.request_id] = n
        req = InferenceRequest(payload.request_id, list(payload.prompt_tokens), sp, status="running")
        req.generated_tokens.append(payload.first_token)
        eng.running.append(req)
        return True

    def step(self):
        return self.engine.step()

    def run_until_done(self) -> Dict[int, InferenceRequest]:
        return self.engine.run_until_done()


# ---- transports ------------------------------------------------------------------------------------------------------------------------------
class InProcessTransport:
    def __init__(self):
        self.queue: List[KVPayload] = []
        self.bytes_moved = 0

    def send(self, payload: KVPayload, dst: int = 0) -> None:
        self.bytes_moved += payload.nbytes
        self.queue.append(payload)

    def recv(self, src: int = 0) -> Optional[KVPayload]:
        return self.queue.pop(0) if self.queue else None


class TorchDistTransport:
    """Header (int64: request id, prompt length, first token, generation budget, layers, kv heads, head dim, dtype code, stop-token count, stop tokens…) then one
    flat K|V tensor.  Sampling temperature / top-k / top-p travel as a small float tensor."""

    _DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
    _HDR = 32

    def __init__(self, group=None, device=None):
        self.group = group
        self.device = device or ("cuda" if dist.is_initialized() and dist.get_backend(group) == "nccl" else "cpu")
        self.bytes_moved = 0

    def send(self, payload: KVPayload, dst: int) -> None:
        sp = payload.sampling_params
        L, n, hk, d = payload.k.shape
        stops = list(sp.stop_token_ids)[: self._HDR - 10]
        hdr = torch.zeros(self._HDR, dtype=torch.int64)
        hdr[:9] = torch.tensor([payload.request_id, n, payload.first_token, sp.num_tokens_to_generate, L, hk, d, self._DT[payload.k.dtype], len(stops)])
        if stops:
            hdr[9 : 9 + len(stops)] = torch.tensor(stops)
        fl = torch.tensor([sp.temperature, float(sp.top_k), sp.top_p], dtype=torch.float32)
        dist.send(hdr.to(self.device), dst, group=self.group)
        dist.send(fl.to(self.device), dst, group=self.group)
        dist.send(torch.tensor(payload.prompt_tokens, dtype=torch.int64, device=self.device), dst, group=self.group)
        flat = torch.cat([payload.k.reshape(-1), payload.v.reshape(-1)]).to(self.device)
        dist.send(flat, dst, group=self.group)
        self.bytes_moved += payload.nbytes

    def recv(self, src: int) -> KVPayload:
        hdr = torch.zeros(self._HDR, dtype=torch.int64, device=self.device)
        dist.recv(hdr, src, group=self.group)
        fl = torch.zeros(3, dtype=torch.float32, device=self.device)
        dist.recv(fl, src, group=self.group)
        rid, n, first, budget, L, hk, d, dtc, ns = (int(x) for x in hdr[:9].tolist())
        toks = torch.zeros(n, dtype=torch.int64, device=self.device)
        dist.recv(toks, src, group=self.group)
        dt = {v: k for k, v in self._DT.items()}[dtc]
        flat = torch.empty(2 * L * n * hk * d, dtype=dt, device=self.device)
        dist.recv(flat, src, group=self.group)
        k, v = flat[: flat.numel() // 2].view(L, n, hk, d), flat[flat.numel() // 2 :].view(L, n, hk, d)
        sp = SamplingParams(temperature=float(fl[0]), top_k=int(fl[1]), top_p=float(fl[2]), num_tokens_to_generate=budget, stop_token_ids=tuple(int(x) for x in hdr[9 : 9 + ns].tolist()))
        self.bytes_moved += 2 * k.numel() * k.element_size()
        return KVPayload(rid, toks.tolist(), first, sp, k, v)
