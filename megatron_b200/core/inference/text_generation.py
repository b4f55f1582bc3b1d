"""Text-level generation API over the token engines (reference ``inference/text_generation_controllers/`` +
``inference/apis/async_llm.py:20`` ``MegatronAsyncLLM`` + ``tools/run_dynamic_text_generation_server.py``).

* ``TextGenerationController`` — tokenize prompts, run an engine, detokenize, stop-string handling.
* ``AsyncLLM``                 — asyncio front end over ``DynamicInferenceEngine``: ``await llm.generate(prompt)`` from many
                                 coroutines; one background task steps the engine (continuous batching) and resolves futures.
* ``TextGenerationServer``     — dependency-free HTTP server (``PUT/POST /api`` with the reference's JSON schema:
                                 ``{"prompts": [...], "tokens_to_generate": n, "temperature": t, "top_k": k, "top_p": p}``) and an
                                 OpenAI-style ``POST /v1/completions``.
"""
from __future__ import annotations

import asyncio
import json
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict, List, Optional, Sequence

from .engine import DynamicInferenceEngine, InferenceRequest, StaticInferenceEngine
from .sampling import SamplingParams


class TextGenerationController:
    def __init__(self, engine, tokenizer):
        self.engine, self.tokenizer = engine, tokenizer

    def tokenize_prompt(self, prompt: str, add_bos: bool = False) -> List[int]:
        ids = list(self.tokenizer.tokenize(prompt))
        bos = getattr(self.tokenizer, "bos", None)
        if add_bos and bos is not None:
            ids = [bos] + ids
        return ids

    def detokenize(self, ids: Sequence[int]) -> str:
        return self.tokenizer.detokenize(list(ids))

    def generate(self, prompts: Sequence[str], params: Optional[SamplingParams] = None, stop: Sequence[str] = ()) -> List[Dict]:
        params = params or SamplingParams()
        toks = [self.tokenize_prompt(p) for p in prompts]
        if isinstance(self.engine, StaticInferenceEngine):
            outs = self.engine.generate(toks, params)
            gen = [o[len(t):] for o, t in zip(outs, toks)]
        else:
            ids = [self.engine.add_request(t, params) for t in toks]
            done = self.engine.run_until_done()
            gen = [done[i].generated_tokens for i in ids]
        results = []
        for p, g in zip(prompts, gen):
            text = self.detokenize(g)
            for s in stop:
                cut = text.find(s)
                if cut >= 0:
                    text = text[:cut]
            results.append({"prompt": p, "text": text, "tokens": list(g)})
        return results


class AsyncLLM:
    """Many concurrent ``await generate(...)`` calls share one continuously-batched engine."""

    def __init__(self, engine: DynamicInferenceEngine, tokenizer=None):
        self.engine, self.tokenizer = engine, tokenizer
        self._futures: Dict[int, asyncio.Future] = {}
        self._task: Optional[asyncio.Task] = None

    async def _loop(self):
        while self._futures:
            finished = self.engine.step()
            for req in finished:
                fut = self._futures.pop(req.request_id, None)
                if fut is not None and not fut.done():
                    fut.set_result(req)
            await asyncio.sleep(0)  # let new requests in between engine steps
        self._task = None

    async def generate_tokens(self, prompt_tokens: List[int], params: Optional[SamplingParams] = None) -> InferenceRequest:
        rid = self.engine.add_request(list(prompt_tokens), params or SamplingParams())
        fut = asyncio.get_running_loop().create_future()
        self._futures[rid] = fut
        if self._task is None:
            self._task = asyncio.create_task(self._loop())
        return await fut

    async def generate(self, prompt: str, params: Optional[SamplingParams] = None) -> str:
        assert self.tokenizer is not None, "text API needs a tokenizer"
        req = await self.generate_tokens(list(self.tokenizer.tokenize(prompt)), params)
        return self.tokenizer.detokenize(req.generated_tokens)


class TextGenerationServer:
    def __init__(self, controller: TextGenerationController, host: str = "127.0.0.1", port: int = 5000):
        self.controller, self.host, self.port = controller, host, port
        self._lock = threading.Lock()
        self.httpd: Optional[ThreadingHTTPServer] = None

    def _params(self, body: dict) -> SamplingParams:
        return SamplingParams(
            temperature=float(body.get("temperature", 1.0)), top_k=int(body.get("top_k", 0)), top_p=float(body.get("top_p", 0.0)),
            num_tokens_to_generate=int(body.get("tokens_to_generate", body.get("max_tokens", 32))),
            return_log_probs=bool(body.get("logprobs", False)), seed=body.get("random_seed", body.get("seed")),
        )

    def handle(self, path: str, body: dict) -> dict:
        if path.rstrip("/") in ("/api", "/generate"):
            prompts = body.get("prompts")
            if not isinstance(prompts, list) or not prompts:
                raise ValueError("prompts must be a non-empty list")
            with self._lock:
                res = self.controller.generate(prompts, self._params(body), stop=body.get("stop", ()))
            return {"text": [r["prompt"] + r["text"] for r in res], "segments": [r["tokens"] for r in res]}
        if path.rstrip("/") == "/v1/completions":
            prompt = body.get("prompt", "")
            prompts = prompt if isinstance(prompt, list) else [prompt]
            with self._lock:
                res = self.controller.generate(prompts, self._params(body), stop=body.get("stop") or ())
            return {"object": "text_completion", "model": body.get("model", "megatron_b200"),
                    "choices": [{"index": i, "text": r["text"], "finish_reason": "length"} for i, r in enumerate(res)]}
        raise KeyError(path)

    def _make_handler(self):
        server = self

        class Handler(BaseHTTPRequestHandler):
            def _do(self):
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    body = json.loads(self.rfile.read(n) or b"{}")
                    out, code = server.handle(self.path, body), 200
                except KeyError:
                    out, code = {"error": "unknown endpoint"}, 404
                except Exception as e:  # bad request
                    out, code = {"error": str(e)}, 400
                data = json.dumps(out).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            do_PUT = do_POST = _do

            def log_message(self, *a):
                pass

        return Handler

    def start(self, background: bool = True):
        self.httpd = ThreadingHTTPServer((self.host, self.port), self._make_handler())
        self.port = self.httpd.server_address[1]
        if background:
            threading.Thread(target=self.httpd.serve_forever, daemon=True).start()
        else:
            self.httpd.serve_forever()

    def stop(self):
        if self.httpd is not None:
            self.httpd.shutdown()
            self.httpd.server_close()
