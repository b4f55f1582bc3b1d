"""Request records, life-cycle events and wire serialisation (reference ``inference/inference_request.py:1-700``)."""
from __future__ import annotations

import base64
import enum
import hashlib
import io
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .engine import InferenceRequest  # noqa: F401

DynamicInferenceRequest = InferenceRequest          # one record type serves both engines here


class Status(enum.Enum):
    """Life cycle of a request; ``InferenceRequest.status`` holds the lower-case name of the state it is in."""

    WAITING_IN_QUEUE = "waiting"
    ACTIVE_AND_GENERATING_TOKENS = "running"
    ACTIVE_BUT_NOT_GENERATING_TOKENS = "prefilling"
    COMPLETED = "finished"
    FAILED = "failed"

    @classmethod
    def of(cls, request: InferenceRequest) -> "Status":
        return cls(request.status)


# ---- wire format: msgpack / JSON friendly dicts ----------------------------------------------------------------------------
_NP_BF16_VIA = torch.int16           # numpy has no bfloat16: ship the bit pattern


def serialize_tensor(t: torch.Tensor) -> Dict[str, Any]:
    """Tensor -> ``{"dtype", "shape", "data"}`` with raw little-endian bytes (no pickle: safe to receive from clients)."""
    t = t.detach().cpu().contiguous()
    raw = t.view(_NP_BF16_VIA).numpy().tobytes() if t.dtype == torch.bfloat16 else t.numpy().tobytes()
    return {"__tensor__": True, "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape), "data": raw}


def deserialize_tensor(d: Dict[str, Any], device=None) -> torch.Tensor:
    dtype = getattr(torch, d["dtype"])
    data = d["data"] if isinstance(d["data"], (bytes, bytearray)) else base64.b64decode(d["data"])
    if dtype == torch.bfloat16:
        t = torch.frombuffer(bytearray(data), dtype=_NP_BF16_VIA).view(torch.bfloat16)
    else:
        t = torch.frombuffer(bytearray(data), dtype=dtype)
    t = t.view(d["shape"])
    return t.to(device) if device is not None else t


def serialize_ndarray(a: np.ndarray) -> Dict[str, Any]:
    a = np.ascontiguousarray(a)
    return {"__ndarray__": True, "dtype": a.dtype.str, "shape": list(a.shape), "data": a.tobytes()}


def deserialize_ndarray(d: Dict[str, Any]) -> np.ndarray:
    data = d["data"] if isinstance(d["data"], (bytes, bytearray)) else base64.b64decode(d["data"])
    return np.frombuffer(data, dtype=np.dtype(d["dtype"])).reshape(d["shape"]).copy()


def unwrap_serialized_tensors(obj: Any, device=None) -> Any:
    """Recursively turn serialised tensors / arrays inside lists and dicts back into objects."""
    if isinstance(obj, dict):
        if obj.get("__tensor__"):
            return deserialize_tensor(obj, device)
        if obj.get("__ndarray__"):
            return deserialize_ndarray(obj)
        return {k: unwrap_serialized_tensors(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(unwrap_serialized_tensors(v, device) for v in obj)
    return obj


def serialize_multimodal_data(data: Optional[Dict[str, Any]]) -> Optional[Dict[str, Any]]:
    """``{"images": [tensor | ndarray | bytes, ...], "image_sizes": ..., ...}`` -> wire form."""
    if data is None:
        return None

    def one(v):
        if torch.is_tensor(v):
            return serialize_tensor(v)
        if isinstance(v, np.ndarray):
            return serialize_ndarray(v)
        if isinstance(v, (list, tuple)):
            return [one(x) for x in v]
        if isinstance(v, dict):
            return {k: one(x) for k, x in v.items()}
        return v
    return one(data)


def resolve_multimodal_data_for_engine(data: Optional[Dict[str, Any]], device=None, dtype: Optional[torch.dtype] = None) -> Optional[Dict[str, Any]]:
    """Wire form -> tensors on the engine's device; arrays become tensors, image tensors are cast to the vision dtype."""
    if data is None:
        return None
    out = unwrap_serialized_tensors(data, device)

    def fix(v):
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
            v = v.to(device) if device is not None else v
        if torch.is_tensor(v) and dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        if isinstance(v, list):
            return [fix(x) for x in v]
        if isinstance(v, dict):
            return {k: fix(x) for k, x in v.items()}
        return v
    return fix(out)


def compute_block_hashes_batched(prompts: Sequence[Sequence[int]], block_size: int, parent: int = 0) -> List[List[int]]:
    """Chained hashes of the FULL blocks of every prompt: ``h_i = H(h_{i-1}, tokens of block i)`` — equal prefixes give equal
    hash chains, which is what the prefix cache and the prefix-aware coordinator key on.  The trailing partial block has no hash."""
    out = []
    for toks in prompts:
        arr = np.asarray(list(toks), dtype=np.int64)
        chain, h = [], parent
        for b in range(len(arr) // block_size):
            m = hashlib.blake2b(digest_size=8)
            m.update(int(h).to_bytes(8, "little", signed=False))
            m.update(arr[b * block_size:(b + 1) * block_size].tobytes())
            h = int.from_bytes(m.digest(), "little")
            chain.append(h)
        out.append(chain)
    return out


# ---- events / records ----------------------------------------------------------------------------------------------------------
class DynamicInferenceEventType(enum.Enum):
    ADD_ENGINE = "add_engine"        # accepted by an engine
    ADD_CONTEXT = "add_context"      # KV pages allocated, prefill scheduled
    GENERATED_TOKEN = "generated_token"
    PAUSE = "pause"                  # evicted under memory pressure, will be re-prefilled
    EVICT = "evict"
    FINISH = "finish"
    FAIL = "fail"
    ERROR_TRANSIENT = "error_transient"
    ERROR_NONTRANSIENT = "error_nontransient"


@dataclass
class DynamicInferenceEvent:
    type: DynamicInferenceEventType
    timestamp: float = field(default_factory=time.time)
    payload: Optional[Any] = None

    def serialize(self) -> Dict[str, Any]:
        p = self.payload
        return {"type": self.type.value, "timestamp": self.timestamp, "payload": serialize_tensor(p) if torch.is_tensor(p) else p}

    @classmethod
    def deserialize(cls, d: Dict[str, Any]) -> "DynamicInferenceEvent":
        return cls(DynamicInferenceEventType(d["type"]), d["timestamp"], unwrap_serialized_tensors(d.get("payload")))


@dataclass
class DynamicInferenceRequestRecord:
    """History of one request across pauses: every (re)admission creates a new engine-side request whose prompt is the original
    prompt plus what had been generated so far; ``merge`` stitches them back into what the client asked for."""
    requests: List[InferenceRequest] = field(default_factory=list)
    events: List[DynamicInferenceEvent] = field(default_factory=list)
    latency: Optional[float] = None

    @classmethod
    def from_request(cls, request: InferenceRequest) -> "DynamicInferenceRequestRecord":
        return cls([request], [DynamicInferenceEvent(DynamicInferenceEventType.ADD_ENGINE)])

    def __getitem__(self, i: int) -> InferenceRequest:
        return self.requests[i]

    @property
    def request_id(self):
        return self.requests[0].request_id

    def checkpoint(self, resumed: InferenceRequest) -> None:
        self.events.append(DynamicInferenceEvent(DynamicInferenceEventType.PAUSE))
        self.requests.append(resumed)

    def add_event(self, type_: DynamicInferenceEventType, payload=None) -> None:
        self.events.append(DynamicInferenceEvent(type_, payload=payload))

    def merge(self) -> InferenceRequest:
        first, last = self.requests[0], self.requests[-1]
        n_prompt = len(first.prompt_tokens)
        full = list(last.prompt_tokens) + list(last.generated_tokens)
        merged = InferenceRequest(first.request_id, list(first.prompt_tokens), first.sampling_params)
        merged.generated_tokens = full[n_prompt:]
        merged.status, merged.arrival_time, merged.finish_time = last.status, first.arrival_time, last.finish_time
        return merged

    def time_to_first_token(self) -> Optional[float]:
        start = next((e.timestamp for e in self.events if e.type is DynamicInferenceEventType.ADD_ENGINE), None)
        tok = next((e.timestamp for e in self.events if e.type is DynamicInferenceEventType.GENERATED_TOKEN), None)
        return None if start is None or tok is None else tok - start


@dataclass
class FinishedRequestRecord:
    """What an engine hands back to the coordinator / client for a completed request."""
    request_id: Any
    prompt_tokens: List[int]
    generated_tokens: List[int]
    generated_text: Optional[str] = None
    logprobs: Optional[List[float]] = None
    finish_reason: str = "stop"
    events: List[DynamicInferenceEvent] = field(default_factory=list)

    def serialize(self) -> Dict[str, Any]:
        return {"request_id": self.request_id, "prompt_tokens": list(self.prompt_tokens), "generated_tokens": list(self.generated_tokens),
                "generated_text": self.generated_text, "logprobs": self.logprobs, "finish_reason": self.finish_reason,
                "events": [e.serialize() for e in self.events]}

    @classmethod
    def deserialize(cls, d: Dict[str, Any]) -> "FinishedRequestRecord":
        return cls(d["request_id"], d["prompt_tokens"], d["generated_tokens"], d.get("generated_text"), d.get("logprobs"), d.get("finish_reason", "stop"),
                   [DynamicInferenceEvent.deserialize(e) for e in d.get("events", [])])


@dataclass
class VLMInferenceRequest:
    """A text request with images: the prompt contains one image placeholder token per image; the engine expands every
    placeholder into ``num_img_embeddings`` positions (``ImageProcessingConfig.embeddings_for``)."""
    request_id: Any
    prompt_tokens: List[int]
    images: List[torch.Tensor] = field(default_factory=list)
    image_token_id: int = -200
    num_img_embeddings: Optional[List[int]] = None
    params: Optional[Any] = None

    def expanded_length(self) -> int:
        n_img = sum(1 for t in self.prompt_tokens if t == self.image_token_id)
        emb = self.num_img_embeddings or [0] * n_img
        assert len(emb) == n_img == len(self.images), f"{n_img} placeholders, {len(self.images)} images, {len(emb)} embedding counts"
        return len(self.prompt_tokens) - n_img + sum(emb)


@dataclass
class DynamicVLMInferenceRequest(VLMInferenceRequest):
    """Dynamic-engine form: position of every image span in the expanded sequence (chunked prefill must not split a span)."""

    def image_spans(self) -> List[tuple]:
        spans, pos, k = [], 0, 0
        emb = self.num_img_embeddings or []
        for t in self.prompt_tokens:
            if t == self.image_token_id:
                spans.append((pos, pos + emb[k]))
                pos += emb[k]
                k += 1
            else:
                pos += 1
        return spans
