"""Text-level generation API over the token engines (reference ``inference/text_generation_controllers/`` +
``inference/apis/async_llm.py:20`` ``MegatronAsyncLLM`` + ``tools/run_dynamic_text_generation_server.py``).

* ``TextGenerationController`` — tokenize prompts, run an engine, detokenize, stop-string handling.
* ``AsyncLLM``                 — asyncio front end over ``DynamicInferenceEngine``: ``await llm.generate(prompt)`` from many
                                 coroutines; one background task steps the engine (continuous batching) and resolves futures.
* ``TextGenerationServer``     — dependency-free HTTP server (``PUT/POST /api`` with the reference's JSON schema:
                                 ``{"prompts": [...], "tokens_to_generate": n, "temperature": t, "top_k": k, "top_p": p}``) and an
                                 OpenAI-style ``POST /v1/completions`` / ``/v1/chat/completions`` (``"stream": true`` → server-sent
                                 events, one chunk per engine step), ``GET /health``, ``GET /v1/models``.
* ``IncrementalDetokenizer``   — turns a growing token list into text deltas without emitting half a UTF-8 character.
* ``AsyncStream``              — per-request async queue fed by the engine loop (``async for delta in llm.generate_stream(...)``).
"""
from __future__ import annotations

import asyncio
import itertools
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


class IncrementalDetokenizer:
    """Reference ``dynamic_text_gen_server/incremental_detokenizer.py``: decode the whole generated prefix each time (tokenizers are not
    prefix-stable: byte-level BPE can end in the middle of a multi-byte character, sentencepiece drops the leading space of a lone piece)
    and emit only the part that is new AND stable — a tail that still decodes to U+FFFD is held back until the next token completes it."""

    def __init__(self, tokenizer):
        self.tok = tokenizer
        self.ids: List[int] = []
        self.emitted = ""

    def add(self, new_ids: Sequence[int], final: bool = False) -> str:
        self.ids.extend(int(i) for i in new_ids)
        text = self.tok.detokenize(list(self.ids))
        if not final:
            stable = len(text)
            while stable > 0 and text[stable - 1] == "\ufffd":
                stable -= 1
            text = text[:stable]
        if not text.startswith(self.emitted):              # the tokenizer rewrote earlier text (rare): resynchronise on the common prefix
            k = 0
            while k < min(len(text), len(self.emitted)) and text[k] == self.emitted[k]:
                k += 1
            self.emitted = self.emitted[:k]
        delta = text[len(self.emitted):]
        self.emitted = text
        return delta


class AsyncStream:
    """Reference ``inference/async_stream.py``: a queue the engine loop ``put``s into and one consumer iterates; ``finish`` ends it."""

    _END = object()

    def __init__(self, request_id: int):
        self.request_id = request_id
        self._q: asyncio.Queue = asyncio.Queue()
        self.finished = False

    def put(self, item) -> None:
        if not self.finished:
            self._q.put_nowait(item)

    def finish(self, exc: Optional[BaseException] = None) -> None:
        if not self.finished:
            self.finished = True
            self._q.put_nowait(exc if exc is not None else self._END)

    def __aiter__(self):
        return self

    async def __anext__(self):
        item = await self._q.get()
        if item is self._END:
            raise StopAsyncIteration
        if isinstance(item, BaseException):
            raise item
        return item


class AsyncLLM:
    """Many concurrent ``await generate(...)`` calls share one continuously-batched engine."""

    def __init__(self, engine: DynamicInferenceEngine, tokenizer=None):
        self.engine, self.tokenizer = engine, tokenizer
        self._futures: Dict[int, asyncio.Future] = {}
        self._streams: Dict[int, AsyncStream] = {}
        self._sent: Dict[int, int] = {}
        self._task: Optional[asyncio.Task] = None

    def _publish(self, finished):
        """Push the tokens produced by the last step to their streams."""
        live = {r.request_id: r for r in list(self.engine.running) + list(finished)}
        for rid, st in list(self._streams.items()):
            req = live.get(rid)
            if req is None:
                continue
            n = self._sent.get(rid, 0)
            if len(req.generated_tokens) > n:
                st.put(list(req.generated_tokens[n:]))
                self._sent[rid] = len(req.generated_tokens)
            if req.status == "finished":
                st.finish()
                self._streams.pop(rid), self._sent.pop(rid, None)

    async def _loop(self):
        while self._futures:
            try:
                finished = self.engine.step()
            except BaseException as e:                     # fail every waiter instead of hanging them
                for fut in self._futures.values():
                    if not fut.done():
                        fut.set_exception(e)
                for st in self._streams.values():
                    st.finish(e)
                self._futures.clear(), self._streams.clear()
                break
            self._publish(finished)
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

    def stream_tokens(self, prompt_tokens: List[int], params: Optional[SamplingParams] = None) -> AsyncStream:
        """→ an ``AsyncStream`` yielding lists of new token ids, one item per engine step that produced any."""
        rid = self.engine.add_request(list(prompt_tokens), params or SamplingParams())
        st = AsyncStream(rid)
        self._streams[rid] = st
        self._futures[rid] = asyncio.get_running_loop().create_future()
        if self._task is None:
            self._task = asyncio.create_task(self._loop())
        return st

    async def generate_stream(self, prompt: str, params: Optional[SamplingParams] = None):
        """``async for text_delta in llm.generate_stream("...")``."""
        assert self.tokenizer is not None, "text API needs a tokenizer"
        det = IncrementalDetokenizer(self.tokenizer)
        async for ids in self.stream_tokens(list(self.tokenizer.tokenize(prompt)), params):
            d = det.add(ids)
            if d:
                yield d
        tail = det.add([], final=True)
        if tail:
            yield tail


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
        if path.rstrip("/") == "/v1/chat/completions":
            prompt = self.render_chat(body.get("messages") or [])
            with self._lock:
                res = self.controller.generate([prompt], self._params(body), stop=body.get("stop") or ())
            r = res[0]
            n_in = len(self.controller.tokenize_prompt(prompt))
            return {"object": "chat.completion", "model": body.get("model", "megatron_b200"),
                    "choices": [{"index": 0, "message": {"role": "assistant", "content": r["text"]},
                                 "finish_reason": "length" if len(r["tokens"]) >= self._params(body).num_tokens_to_generate else "stop"}],
                    "usage": {"prompt_tokens": n_in, "completion_tokens": len(r["tokens"]), "total_tokens": n_in + len(r["tokens"])}}
        raise KeyError(path)

    def render_chat(self, messages: List[dict]) -> str:
        """The tokenizer's own chat template when it has one (HF tokenizers), else a plain ``role: content`` transcript."""
        tok = self.controller.tokenizer
        inner = getattr(tok, "tokenizer", tok)
        if hasattr(inner, "apply_chat_template") and getattr(inner, "chat_template", None):
            return inner.apply_chat_template(messages, tokenize=False, add_generation_prompt=True)
        for m in messages:
            if not isinstance(m, dict) or "role" not in m or "content" not in m:
                raise ValueError("each message needs 'role' and 'content'")
        return "".join(f"{m['role']}: {m['content']}\n" for m in messages) + "assistant: "

    def handle_get(self, path: str) -> dict:
        if path.rstrip("/") == "/health":
            eng = self.controller.engine
            return {"status": "ok", "running": len(getattr(eng, "running", [])), "waiting": len(getattr(eng, "waiting", []))}
        if path.rstrip("/") == "/v1/models":
            return {"object": "list", "data": [{"id": "megatron_b200", "object": "model"}]}
        raise KeyError(path)

    def stream(self, path: str, body: dict):
        """Generator of SSE ``data:`` payloads (dicts; ``None`` = ``[DONE]``) for a ``"stream": true`` request."""
        chat = path.rstrip("/") == "/v1/chat/completions"
        if not chat and path.rstrip("/") != "/v1/completions":
            raise KeyError(path)
        prompt = self.render_chat(body.get("messages") or []) if chat else body.get("prompt", "")
        eng = self.controller.engine
        if not isinstance(eng, DynamicInferenceEngine):
            raise ValueError("streaming needs the dynamic engine")
        det = IncrementalDetokenizer(self.controller.tokenizer)
        obj = "chat.completion.chunk" if chat else "text_completion"

        def chunk(text, finish=None):
            c = {"index": 0, "finish_reason": finish}
            c.update({"delta": ({"content": text} if text else {})} if chat else {"text": text})
            return {"object": obj, "model": body.get("model", "megatron_b200"), "choices": [c]}

        with self._lock:
            rid = eng.add_request(self.controller.tokenize_prompt(prompt), self._params(body))
            sent, req = 0, None
            while req is None or req.status != "finished":
                done = eng.step()
                req = next((r for r in list(eng.running) + list(done) if r.request_id == rid), None) or eng.finished[rid]
                if len(req.generated_tokens) > sent:
                    d = det.add(req.generated_tokens[sent:])
                    sent = len(req.generated_tokens)
                    if d:
                        yield chunk(d)
            tail = det.add([], final=True)
            yield chunk(tail, "length" if len(req.generated_tokens) >= req.sampling_params.num_tokens_to_generate else "stop")
        yield None

    def _make_handler(self):
        server = self

        class Handler(BaseHTTPRequestHandler):
            def _do(self):
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    body = json.loads(self.rfile.read(n) or b"{}")
                    if body.get("stream"):
                        gen = server.stream(self.path, body)
                        first = next(gen)                  # errors surface before the 200 is sent
                        self.send_response(200)
                        self.send_header("Content-Type", "text/event-stream")
                        self.send_header("Cache-Control", "no-cache")
                        self.end_headers()
                        for ev in itertools.chain([first], gen):
                            self.wfile.write(b"data: " + (json.dumps(ev).encode() if ev is not None else b"[DONE]") + b"\n\n")
                            self.wfile.flush()
                        return
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

            def do_GET(self):
                try:
                    out, code = server.handle_get(self.path), 200
                except KeyError:
                    out, code = {"error": "unknown endpoint"}, 404
                data = json.dumps(out).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

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
