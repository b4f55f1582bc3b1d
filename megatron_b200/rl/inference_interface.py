"""How the RL loop talks to a generator (reference ``megatron/rl/inference/``: ``InferenceInterface``, ``ReturnsRaw``, chat / completion request types, the
Megatron-local and remote (OpenAI-compatible HTTP) back-ends).

``LocalEngineInference`` drives this repo's engines in process (colocated training + generation — weights are shared, so "refit" is free);
``RemoteHTTPInference`` posts to ``tools/run_text_generation_server.py`` (``/api``), for generation pools that live elsewhere; refitting those means pushing
weights through ``core/resharding`` or a checkpoint, which the caller schedules with ``WeightRefitter``."""
from __future__ import annotations

import json
import urllib.request
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import torch

from ..core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
from ..core.inference.sampling import SamplingParams


@dataclass
class InferenceRequest:
    prompt_tokens: List[int]
    sampling: SamplingParams = field(default_factory=SamplingParams)
    n: int = 1                                   # samples per prompt (a GRPO group)


@dataclass
class InferenceResponse:
    prompt_tokens: List[int]
    completions: List[List[int]]
    logprobs: Optional[List[List[float]]] = None
    policy_version: int = 0


class InferenceInterface(ABC):
    policy_version: int = 0

    @abstractmethod
    def generate(self, requests: List[InferenceRequest]) -> List[InferenceResponse]:
        ...

    def set_policy_version(self, v: int) -> None:
        self.policy_version = v


class LocalEngineInference(InferenceInterface):
    def __init__(self, model, vocab_size: Optional[int] = None, dynamic: bool = False, max_sequence_length: int = 2048, **engine_kw):
        self.model = model
        self.dynamic = dynamic
        self.engine = (DynamicInferenceEngine(model, vocab_size=vocab_size, **engine_kw) if dynamic
                       else StaticInferenceEngine(model, max_batch_size=engine_kw.get("max_batch_size", 256), max_sequence_length=max_sequence_length, vocab_size=vocab_size))

    @torch.no_grad()
    def generate(self, requests: List[InferenceRequest]) -> List[InferenceResponse]:
        was_training = self.model.training
        self.model.eval()
        try:
            out: List[InferenceResponse] = []
            if self.dynamic:
                ids = [[self.engine.add_request(r.prompt_tokens, r.sampling) for _ in range(r.n)] for r in requests]
                done = self.engine.run_until_done()
                for r, group in zip(requests, ids):
                    out.append(InferenceResponse(r.prompt_tokens, [done[i].generated_tokens for i in group], [done[i].log_probs for i in group] if r.sampling.return_log_probs else None,
                                                 self.policy_version))
                return out
            for r in requests:      # the static engine batches equal-length prompts itself: a group is one batch
                gens = self.engine.generate([r.prompt_tokens] * r.n, r.sampling)
                out.append(InferenceResponse(r.prompt_tokens, gens, None, self.policy_version))
            return out
        finally:
            self.model.train(was_training)


class RemoteHTTPInference(InferenceInterface):
    def __init__(self, url: str, tokenizer=None, timeout: float = 600.0):
        self.url, self.tokenizer, self.timeout = url.rstrip("/"), tokenizer, timeout

    def generate(self, requests: List[InferenceRequest]) -> List[InferenceResponse]:
        out = []
        for r in requests:
            body = {"prompt_tokens": [r.prompt_tokens] * r.n, "tokens_to_generate": r.sampling.num_tokens_to_generate, "temperature": r.sampling.temperature,
                    "top_k": r.sampling.top_k, "top_p": r.sampling.top_p}
            req = urllib.request.Request(self.url + "/api", data=json.dumps(body).encode(), method="PUT", headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=self.timeout) as resp:
                data = json.loads(resp.read())
            out.append(InferenceResponse(r.prompt_tokens, data.get("tokens") or data.get("completions") or [], data.get("logprobs"), self.policy_version))
        return out


class WeightRefitter:
    """Push the trainer's weights into a generation model that does not share storage with it (reference: ``resharding/refit.py::swap_model_weights``).
    Same layout and one process: parameter-wise copy.  Different layouts, or a distributed job: the refit planner + a copy service
    (``method`` = ``--refit-method``: ``nccl`` | ``gloo`` | ``nvlink``; the reference's ``nvshmem`` resolves to ``nvlink``)."""

    def __init__(self, src_model, dst_model, method: str = "nccl", group=None, src_rank_offset: int = 0, dst_rank_offset: int = 0):
        self.src, self.dst = src_model, dst_model
        self.method, self.group, self.offsets = method, group, (src_rank_offset, dst_rank_offset)
        self.version = 0

    def _same_layout(self) -> bool:
        if self.src is None or self.dst is None:
            return False
        src = dict(self.src.named_parameters())
        dst = dict(self.dst.named_parameters())
        return src.keys() == dst.keys() and all(src[n].shape == dst[n].shape for n in src)

    @torch.no_grad()
    def refit(self) -> int:
        import torch.distributed as dist
        if self._same_layout() and not (dist.is_initialized() and dist.get_world_size(self.group) > 1 and any(self.offsets)):
            src = dict(self.src.named_parameters())
            for name, p in self.dst.named_parameters():
                p.copy_(src[name])
        else:
            from ..core.resharding import swap_model_weights
            method = self.method if (dist.is_initialized() and dist.get_backend(self.group) != "gloo") else "gloo"
            swap_model_weights(self.src, self.dst, method, group=self.group, src_rank_offset=self.offsets[0], dst_rank_offset=self.offsets[1])
        self.version += 1
        return self.version
