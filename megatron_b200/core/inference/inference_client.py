"""HTTP client of the text-generation server (reference ``inference/inference_client.py`` talks ZMQ to the coordinator; the
serving surface here is HTTP, so the client is too).  Blocking calls plus an SSE reader for streamed completions."""
from __future__ import annotations

import json
import urllib.request
from typing import Dict, Iterator, List, Optional, Sequence


class InferenceClient:
    def __init__(self, host: str = "127.0.0.1", port: int = 5000, timeout: float = 600.0):
        self.base, self.timeout = f"http://{host}:{port}", timeout

    def _call(self, path: str, body: Optional[dict] = None, method: str = "POST"):
        data = json.dumps(body).encode() if body is not None else None
        req = urllib.request.Request(self.base + path, data=data, headers={"Content-Type": "application/json"}, method=method)
        return urllib.request.urlopen(req, timeout=self.timeout)

    def health(self) -> dict:
        return json.loads(self._call("/health", method="GET").read())

    def generate(self, prompts: Sequence[str], tokens_to_generate: int = 32, **sampling) -> List[str]:
        out = json.loads(self._call("/api", {"prompts": list(prompts), "tokens_to_generate": tokens_to_generate, **sampling}, "PUT").read())
        return out["text"]

    def completions(self, prompt: str, max_tokens: int = 32, **sampling) -> dict:
        return json.loads(self._call("/v1/completions", {"prompt": prompt, "max_tokens": max_tokens, **sampling}).read())

    def chat(self, messages: List[Dict[str, str]], max_tokens: int = 32, **sampling) -> dict:
        return json.loads(self._call("/v1/chat/completions", {"messages": messages, "max_tokens": max_tokens, **sampling}).read())

    def stream(self, prompt_or_messages, max_tokens: int = 32, **sampling) -> Iterator[str]:
        """Text deltas of a streamed completion (``str`` prompt) or chat completion (``list`` of messages)."""
        chat = not isinstance(prompt_or_messages, str)
        body = {"max_tokens": max_tokens, "stream": True, **sampling}
        body["messages" if chat else "prompt"] = prompt_or_messages
        with self._call("/v1/chat/completions" if chat else "/v1/completions", body) as r:
            for raw in r:
                line = raw.decode().strip()
                if not line.startswith("data: "):
                    continue
                if line == "data: [DONE]":
                    return
                c = json.loads(line[6:])["choices"][0]
                d = c.get("delta", {}).get("content") if chat else c.get("text")
                if d:
                    yield d
