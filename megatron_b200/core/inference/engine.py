"""Inference engines (reference ``inference/engines``: ``StaticInferenceEngine``, ``DynamicInferenceEngine``).

* ``StaticInferenceEngine``  — pads a batch of prompts, prefill once, decode step by step with the static
  KV cache of ``InferenceParams``.
* ``DynamicInferenceEngine`` — continuous batching: requests enter and leave between decode steps, KV lives in
  a paged cache (``kv_cache.PagedKVCache``), each step runs prefill for newly admitted requests and one decode
  token for the running ones; finished requests release their blocks immediately.
"""
from __future__ import annotations

import itertools
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional

import torch

from ..inference_params import InferenceParams
from .kv_cache import PagedKVCache
from .sampling import SamplingParams, sample


@dataclass
class InferenceRequest:
    request_id: int
    prompt_tokens: List[int]
    sampling_params: SamplingParams = field(default_factory=SamplingParams)
    generated_tokens: List[int] = field(default_factory=list)
    log_probs: List[float] = field(default_factory=list)
    status: str = "waiting"  # waiting | prefilling (chunked prefill under way) | running | finished
    prefill_pos: int = 0     # prompt tokens whose K/V are in the cache
    arrival_time: float = field(default_factory=time.time)
    first_token_time: Optional[float] = None
    finish_time: Optional[float] = None

    @property
    def ttft(self):
        return None if self.first_token_time is None else self.first_token_time - self.arrival_time


class StaticInferenceEngine:
    def __init__(self, model, tokenizer=None, max_batch_size: int = 8, max_sequence_length: int = 2048, vocab_size: Optional[int] = None):
        self.model, self.tokenizer = model, tokenizer
        self.max_batch_size, self.max_sequence_length = max_batch_size, max_sequence_length
        self.vocab_size = vocab_size

    @torch.no_grad()
    def generate(self, prompts: List[List[int]], params: Optional[SamplingParams] = None) -> List[List[int]]:
        params = params or SamplingParams()
        self.model.eval()
        dev = next(self.model.parameters()).device
        out: List[List[int]] = [None] * len(prompts)
        # group equal-length prompts so no padding enters the causal cache
        by_len: Dict[int, List[int]] = {}
        for i, p in enumerate(prompts):
            by_len.setdefault(len(p), []).append(i)
        gen = torch.Generator(device=dev)
        if params.seed is not None:
            gen.manual_seed(params.seed)
        for plen, idxs in by_len.items():
            for lo in range(0, len(idxs), self.max_batch_size):
                group = idxs[lo : lo + self.max_batch_size]
                toks = torch.tensor([prompts[i] for i in group], device=dev)
                ctx = InferenceParams(len(group), min(self.max_sequence_length, plen + params.num_tokens_to_generate))
                pos = torch.arange(plen, device=dev)[None].expand(len(group), -1)
                logits = self.model(toks, pos, None, inference_context=ctx)[:, -1]
                ctx.sequence_len_offset = plen
                done = torch.zeros(len(group), dtype=torch.bool, device=dev)
                gens = [[] for _ in group]
                for step in range(params.num_tokens_to_generate):
                    nxt = sample(logits.float(), params.temperature, params.top_k, params.top_p, gen, self.vocab_size)
                    for j, t in enumerate(nxt.tolist()):
                        if not done[j]:
                            gens[j].append(t)
                            if t in params.stop_token_ids:
                                done[j] = True
                    if bool(done.all()) or step == params.num_tokens_to_generate - 1:
                        break
                    p1 = torch.full((len(group), 1), plen + step, device=dev)
                    logits = self.model(nxt[:, None], p1, None, inference_context=ctx)[:, -1]
                    ctx.sequence_len_offset += 1
                for j, i in enumerate(group):
                    out[i] = gens[j]
        return out


class DynamicInferenceEngine:
    """Continuous batching over a paged KV cache.  The model is driven one request-slice at a time through its
    attention layers' ``paged`` hook (``set_paged_context``), which keeps this engine independent of the
    attention kernel: prefill = full causal attention on the prompt, decode = 1 query against the gathered cache."""

    def __init__(self, model, num_blocks: int = 256, block_size: int = 16, max_running: int = 16, vocab_size: Optional[int] = None,
                 batched_decode: Optional[bool] = None, enable_prefix_caching: bool = False, decode_batch_buckets: Optional[List[int]] = None,
                 enable_cuda_graphs: bool = False, max_prefill_tokens_per_step: Optional[int] = None):
        self.model = model
        # Chunked prefill (reference dynamic_context.py / dynamic_engine.py: ``max_tokens`` per step + chunked-prefill requests): a step spends at most this many
        # prompt tokens on prefill, so a long prompt is spread over several steps while the running requests keep producing one token per step (bounded
        # inter-token latency).  A partially prefilled request holds its pages, sits in ``running`` with status "prefilling" and emits nothing yet.
        self.max_prefill_tokens_per_step = max_prefill_tokens_per_step
        self.prefill_chunks = 0
        # CUDA-graphed decode (reference dynamic_engine.py: cuda-graph batch-size buckets): one captured graph per (batch bucket, attended-length bucket, table
        # width); a step copies the block table / positions / last tokens into the graph's static tensors and replays it — the ~500 kernel launches of an
        # eager decode step (16 ms of host time on an 8B model) collapse into one.  Needs static shapes, hence the buckets.
        self.enable_cuda_graphs = enable_cuda_graphs
        self._graphs: Dict[tuple, tuple] = {}
        self.graph_replays = 0
        if enable_cuda_graphs and not decode_batch_buckets:
            decode_batch_buckets = [b for b in (1, 2, 4, 8, 16, 32, 64, 128, 256) if b <= max(1, max_running)] or [max_running]
            if decode_batch_buckets[-1] < max_running:
                decode_batch_buckets.append(max_running)
        # one forward for ALL running requests' next token (block-table attention); models whose attention is not the standard
        # ``Attention`` (MLA latent cache, Mamba state) keep the per-request path
        if batched_decode is None:
            from ..transformer.attention import SelfAttention

            batched_decode = all(isinstance(getattr(l, "self_attention", None), SelfAttention) for l in model.decoder.layers)
        self.batched_decode = batched_decode
        self.decode_forwards = 0
        # static decode shapes (CUDA-graph buckets): the batch is padded up to the next bucket, the attended length to a block multiple
        self.decode_batch_buckets = sorted(decode_batch_buckets) if decode_batch_buckets else None
        self.decode_shapes_seen = set()
        self.prefill_tokens = 0
        cfg = model.config
        dev = next(model.parameters()).device
        dt = next(model.parameters()).dtype
        tp = cfg.tensor_model_parallel_size
        ol = getattr(model, "output_layer", None)
        if tp > 1 and ol is not None and getattr(ol, "gather_output", True) is False:
            # a training-style model returns vocabulary-PARALLEL logits: every TP rank would sample from its own shard.  Serving needs the gathered distribution.
            ol.gather_output = True
            model.parallel_output = False
        self.cache = PagedKVCache(cfg.num_layers, num_blocks, block_size, max(cfg.num_query_groups // tp, 1), cfg.kv_channels, dt, dev, enable_prefix_caching)
        self.waiting: Deque[InferenceRequest] = deque()
        self.running: List[InferenceRequest] = []
        self.finished: Dict[int, InferenceRequest] = {}
        self.max_running = max_running
        self.vocab_size = vocab_size
        self._ids = itertools.count()
        self.device = dev
        self.steps = 0

    def add_request(self, prompt_tokens: List[int], sampling_params: Optional[SamplingParams] = None) -> int:
        rid = next(self._ids)
        self.waiting.append(InferenceRequest(rid, list(prompt_tokens), sampling_params or SamplingParams()))
        return rid

    def has_unfinished(self) -> bool:
        return bool(self.waiting or self.running)

    @torch.no_grad()
    def _forward_request(self, req: InferenceRequest, tokens: List[int], start: int) -> torch.Tensor:
        """Run ``tokens`` (positions start..) of one request through the model using the paged cache."""
        from ..transformer.attention import Attention

        cache, rid = self.cache, req.request_id
        toks = torch.tensor([tokens], device=self.device)
        pos = torch.arange(start, start + len(tokens), device=self.device)[None]
        if start == 0 and self.batched_decode:
            # prompt without a cached prefix: plain causal attention over the prompt (flash kernel); every layer appends its K/V to the request's pages itself
            # — no per-layer contiguous K/V staging buffers, no gather of the (empty) history
            from .kv_cache import PagedPrefillContext

            ctx = PagedPrefillContext(cache, rid, len(tokens), [l.self_attention.layer_number for l in self.model.decoder.layers])
            logits = self.model(toks, pos, None, inference_context=ctx)
            cache.lengths[rid] = len(tokens)
            return logits[0, -1]

        class _Ctx:  # duck-typed inference context understood by Attention._adjust_key_value_for_inference
            max_sequence_length = start + len(tokens)
            max_batch_size = 1
            sequence_len_offset = start
            batch_size_offset = 0
            key_value_memory_dict: Dict = {}

        ctx = _Ctx()
        ctx.key_value_memory_dict = {}
        total = start + len(tokens)
        # materialise this request's cache prefix per layer as a contiguous view the static path understands
        for li, layer in enumerate(self.model.decoder.layers):
            kbuf = torch.zeros(total, 1, cache.k.shape[3], cache.k.shape[4], dtype=cache.k.dtype, device=self.device)
            vbuf = torch.zeros_like(kbuf)
            if start > 0:
                k, v = cache.gather(li, rid, start)
                kbuf[:start, 0], vbuf[:start, 0] = k, v
            ctx.key_value_memory_dict[layer.self_attention.layer_number] = (kbuf, vbuf)
        logits = self.model(toks, pos, None, inference_context=ctx)
        for li, layer in enumerate(self.model.decoder.layers):
            kbuf, vbuf = ctx.key_value_memory_dict[layer.self_attention.layer_number]
            cache.append(li, rid, kbuf[start:total, 0], vbuf[start:total, 0], start)
        cache.lengths[rid] = total
        return logits[0, -1]

    @torch.no_grad()
    def _forward_decode_batch(self, reqs: List[InferenceRequest]) -> torch.Tensor:
        """→ next-token logits ``[B, vocab]`` for ``reqs`` from a single forward."""
        from .kv_cache import BatchedDecodeContext

        rids = [r.request_id for r in reqs]
        pad_to = None
        if self.decode_batch_buckets:
            pad_to = next((b for b in self.decode_batch_buckets if b >= len(rids)), len(rids))
        ctx = BatchedDecodeContext(self.cache, rids, [l.self_attention.layer_number for l in self.model.decoder.layers], pad_to,
                                   self.cache.block_size * 4 if self.decode_batch_buckets else 1)
        last = [[r.generated_tokens[-1]] for r in reqs] + [[0]] * (ctx.lengths.numel() - len(reqs))
        toks = torch.tensor(last, device=self.device)
        self.decode_shapes_seen.add((toks.shape[0], ctx.max_len))
        if self.enable_cuda_graphs and toks.is_cuda:
            logits = self._graphed_decode(toks, ctx)[: len(reqs)]
        else:
            logits = self.model(toks, ctx.lengths[:, None], None, inference_context=ctx)[: len(reqs)]   # [B, 1, vocab]
        for r in rids:
            self.cache.lengths[r] += 1
        self.decode_forwards += 1
        return logits[:, -1]

    def _graphed_decode(self, toks: torch.Tensor, ctx) -> torch.Tensor:
        key = (toks.shape[0], ctx.max_len, ctx.block_table.shape[1])
        entry = self._graphs.get(key)
        if entry is None:
            s_toks, s_ctx = toks.clone(), ctx            # this step's context becomes the graph's static context
            # eager warm-up on a side stream (autotuning, lazy initialisation) — also the correct result for this step
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.model(s_toks, s_ctx.lengths[:, None], None, inference_context=s_ctx)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.model(s_toks, s_ctx.lengths[:, None], None, inference_context=s_ctx)
            self._graphs[key] = entry = (graph, s_toks, s_ctx, out)
            # the warm-up and the capture both appended this step's K/V at the same slot (idempotent); replay once so `out` holds this step's logits
            graph.replay()
            self.graph_replays += 1
            return out.clone()
        graph, s_toks, s_ctx, out = entry
        s_toks.copy_(toks)
        s_ctx.copy_from(ctx)
        graph.replay()
        self.graph_replays += 1
        return out.clone()

    @torch.no_grad()
    def step(self) -> List[InferenceRequest]:
        """Admit what fits, run one token for every running request, retire the finished ones."""
        self.model.eval()
        newly_finished: List[InferenceRequest] = []
        admitted: List[InferenceRequest] = []
        budget = self.max_prefill_tokens_per_step if self.max_prefill_tokens_per_step else float("inf")

        def prefill_some(req: InferenceRequest) -> None:
            """Run the next chunk of ``req``'s prompt (all of it when the budget allows); the first token is emitted with the last chunk."""
            nonlocal budget
            n_prompt = len(req.prompt_tokens)
            take = int(min(n_prompt - req.prefill_pos, budget))
            if take <= 0:
                return
            start = req.prefill_pos
            logits = self._forward_request(req, req.prompt_tokens[start:start + take], start)
            req.prefill_pos += take
            budget -= take
            self.prefill_tokens += take
            self.prefill_chunks += 1
            if req.prefill_pos >= n_prompt:
                req.status = "running"
                self.cache.register_prefix(req.request_id, req.prompt_tokens)
                self._emit(req, logits)
            admitted.append(req)                                       # no decode for it in this step

        for req in self.running:                                       # 1. chunked prefills under way, oldest first
            if req.status == "prefilling" and budget > 0:
                prefill_some(req)
        while self.waiting and len(self.running) < self.max_running and budget > 0:      # 2. admissions
            req = self.waiting[0]
            need = len(req.prompt_tokens) + req.sampling_params.num_tokens_to_generate
            if not self.cache.can_admit(need) or not self.cache.add_request(req.request_id, len(req.prompt_tokens), req.prompt_tokens):
                break
            self.waiting.popleft()
            req.status = "prefilling"
            req.prefill_pos = self.cache.prefix_hit_tokens.get(req.request_id, 0)       # leading tokens whose K/V are already cached
            self.running.append(req)
            prefill_some(req)
        ready = []
        for req in list(self.running):
            if req.status in ("finished", "prefilling") or req in admitted:
                continue
            if len(req.generated_tokens) >= req.sampling_params.num_tokens_to_generate:
                continue
            cur = self.cache.lengths[req.request_id]
            if not self.cache.ensure_capacity(req.request_id, cur + 1):
                continue  # out of blocks this step; try again after others finish
            ready.append(req)
        if ready and self.batched_decode:
            self._emit_batch(ready, self._forward_decode_batch(ready))
        else:
            for req in ready:
                logits = self._forward_request(req, [req.generated_tokens[-1]], self.cache.lengths[req.request_id])
                self._emit(req, logits)
        for req in list(self.running):
            sp = req.sampling_params
            if req.status == "prefilling":
                continue
            if len(req.generated_tokens) >= sp.num_tokens_to_generate or (req.generated_tokens and req.generated_tokens[-1] in sp.stop_token_ids):
                req.status, req.finish_time = "finished", time.time()
                self.cache.release(req.request_id)
                self.running.remove(req)
                self.finished[req.request_id] = req
                newly_finished.append(req)
        self.steps += 1
        return newly_finished

    def _emit(self, req: InferenceRequest, logits: torch.Tensor):
        sp = req.sampling_params
        tok = int(sample(logits.float()[None], sp.temperature, sp.top_k, sp.top_p, None, self.vocab_size)[0])
        if req.first_token_time is None:
            req.first_token_time = time.time()
        if sp.return_log_probs:
            req.log_probs.append(float(torch.log_softmax(logits.float(), -1)[tok]))
        req.generated_tokens.append(tok)

    def _emit_batch(self, reqs: List[InferenceRequest], logits: torch.Tensor):
        """Sample the next token of every request of a decode step with ONE device→host transfer per group of equal sampling parameters (the per-request
        ``int(...)`` of ``_emit`` costs a host sync per request and token)."""
        groups: Dict[tuple, List[int]] = {}
        for i, r in enumerate(reqs):
            sp = r.sampling_params
            groups.setdefault((sp.temperature, sp.top_k, sp.top_p, sp.return_log_probs), []).append(i)
        now = None
        for (temp, top_k, top_p, want_lp), idxs in groups.items():
            lg = logits[idxs].float() if len(idxs) != len(reqs) else logits.float()
            toks = sample(lg, temp, top_k, top_p, None, self.vocab_size)
            lps = torch.log_softmax(lg, -1).gather(1, toks[:, None]).squeeze(1).tolist() if want_lp else None
            for j, (i, t) in enumerate(zip(idxs, toks.tolist())):
                r = reqs[i]
                if r.first_token_time is None:
                    now = now or time.time()
                    r.first_token_time = now
                if lps is not None:
                    r.log_probs.append(lps[j])
                r.generated_tokens.append(int(t))

    def run_until_done(self) -> Dict[int, InferenceRequest]:
        while self.has_unfinished():
            self.step()
        return self.finished
