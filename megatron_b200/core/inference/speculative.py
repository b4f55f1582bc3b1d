"""Speculative decoding (reference ``inference/text_generation_controllers`` speculative path + ``mtp_utils_triton.py``: verify / rewind).

A cheap *draft* proposes ``k`` tokens autoregressively; the target model scores all ``k + 1`` positions in ONE forward on its KV cache; the longest
accepted prefix is kept, the first rejected position is replaced by the target's own token, and both caches are rewound to the accepted length.
The output distribution is exactly the target's: greedy decoding accepts a draft token iff it equals the target argmax; sampling uses the standard
rejection rule (accept with probability ``min(1, p_target / p_draft)``, resample the first rejection from ``normalize(max(0, p_t - p_d))``).

Rewinding a static cache is free: ``InferenceParams.sequence_len_offset`` is simply set back — stale rows beyond it are overwritten by the next
forward and never attended to (causal mask on absolute positions).  Decode is launch- and bandwidth-bound on a B200 (one token reads every
weight once), so verifying ``k + 1`` tokens costs about the same as one: acceptance rate ≈ speed-up.

The draft can be any callable with the model interface: a smaller model, or the target's own multi-token-prediction heads wrapped in ``MTPDraft``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from ..inference_params import InferenceParams
from .sampling import SamplingParams


@dataclass
class SpeculativeStats:
    proposed: int = 0
    accepted: int = 0
    target_forwards: int = 0
    draft_forwards: int = 0
    per_step_accepted: List[int] = field(default_factory=list)

    @property
    def acceptance_rate(self) -> float:
        return self.accepted / max(self.proposed, 1)

    @property
    def tokens_per_target_forward(self) -> float:
        return (self.accepted + len(self.per_step_accepted)) / max(self.target_forwards, 1)


def _probs(logits: torch.Tensor, p: SamplingParams, vocab_size: Optional[int]) -> torch.Tensor:
    logits = logits.float()
    if vocab_size is not None and vocab_size < logits.shape[-1]:
        logits = logits.clone()
        logits[..., vocab_size:] = float("-inf")
    if p.temperature == 0.0 or p.top_k == 1:
        return torch.nn.functional.one_hot(logits.argmax(-1), logits.shape[-1]).float()
    logits = logits / max(p.temperature, 1e-6)
    if p.top_k > 0:
        kth = torch.topk(logits, min(p.top_k, logits.shape[-1]), dim=-1).values[..., -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    return torch.softmax(logits, dim=-1)


def verify_draft_tokens(draft_tokens: torch.Tensor, draft_probs: torch.Tensor, target_probs: torch.Tensor, generator: Optional[torch.Generator] = None):
    """One sequence.  draft_tokens [k]; draft_probs [k, v] (the distributions the draft sampled from); target_probs [k + 1, v].
    → (n_accepted, next_token): the token that follows the accepted prefix (a correction, or the bonus token when everything was accepted)."""
    k = draft_tokens.shape[0]
    idx = torch.arange(k, device=draft_tokens.device)
    pt = target_probs[idx, draft_tokens]
    pd = draft_probs[idx, draft_tokens].clamp(min=1e-20)
    u = torch.rand(k, device=draft_tokens.device, generator=generator)
    rejected = (u >= (pt / pd).clamp(max=1.0)).nonzero()
    n = int(rejected[0]) if rejected.numel() else k
    if n == k:
        dist = target_probs[k]
    else:
        dist = (target_probs[n] - draft_probs[n]).clamp(min=0)
        dist = dist / dist.sum() if dist.sum() > 0 else target_probs[n]
    nxt = torch.multinomial(dist, 1, generator=generator) if (dist > 0).sum() > 1 else dist.argmax().view(1)
    return n, int(nxt)


def verify_draft_tokens_batched(draft_tokens: torch.Tensor, draft_probs: Optional[torch.Tensor], target_probs: torch.Tensor, u_accept: Optional[torch.Tensor] = None,
                                u_sample: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
    """A whole decode batch at once, no host sync (reference ``text_generation_controllers/mtp_utils_triton.py``: verify + rewind counts).

    ``draft_tokens [B, k]``; ``draft_probs [B, k, V]`` or ``None`` (greedy / MTP drafts: one-hot on the draft token); ``target_probs [B, k + 1, V]``;
    ``u_accept [B, k]`` / ``u_sample [B]`` uniforms (drawn here when omitted).  → ``(n_accepted [B], next_token [B])`` int64 on the device; the caller rewinds
    each KV cache to ``len + n_accepted``.  Acceptance: ``u < min(1, p_t / p_d)``; the follow-up token is drawn from ``normalise(max(p_t - p_d, 0))`` (or the
    bonus row) by an inverse-CDF walk in vocabulary order, so CPU and GPU agree given the same uniforms.  CUDA: one block per sequence (``csrc/misc_kernels.cu``)."""
    B, k = draft_tokens.shape
    dev = draft_tokens.device
    if u_accept is None:
        u_accept = torch.rand(B, k, device=dev, generator=generator)
    if u_sample is None:
        u_sample = torch.rand(B, device=dev, generator=generator)
    from ... import ops

    if draft_tokens.is_cuda and ops.has_ext() and hasattr(ops.ext(), "spec_verify"):
        n, nxt = ops.ext().spec_verify(draft_tokens.long().contiguous(), None if draft_probs is None else draft_probs.float().contiguous(), target_probs.float().contiguous(),
                                       u_accept.float().contiguous(), u_sample.float().contiguous())
        ops._count()
        return n, nxt
    tp = target_probs.float()
    V = tp.shape[-1]
    dp = draft_probs.float() if draft_probs is not None else torch.nn.functional.one_hot(draft_tokens.long(), V).float()
    idx = draft_tokens.long().unsqueeze(-1)
    pt = tp[:, :k].gather(-1, idx).squeeze(-1)
    pd = dp.gather(-1, idx).squeeze(-1).clamp(min=1e-20) if draft_probs is not None else torch.ones_like(pt)
    accept = u_accept < (pt / pd).clamp(max=1.0)
    n = torch.cumprod(accept.to(torch.int64), dim=1).sum(1)                                   # length of the accepted prefix
    rows = torch.arange(B, device=dev)
    trow = tp[rows, n]
    drow = torch.where((n < k).unsqueeze(-1), dp[rows, n.clamp(max=k - 1)] if k > 0 else torch.zeros_like(trow), torch.zeros_like(trow))
    resid = (trow - drow).clamp(min=0)
    dead = resid.sum(-1, keepdim=True) <= 0
    dist = torch.where(dead, trow, resid)
    cdf = dist.cumsum(-1)
    thr = (u_sample.float() * cdf[:, -1]).unsqueeze(-1)
    hit = (cdf > thr) & (dist > 0)
    any_hit = hit.any(-1)
    first = hit.float().argmax(-1)
    last_nz = (V - 1) - (dist > 0).flip(-1).float().argmax(-1)
    return n, torch.where(any_hit, first, last_nz)


class SpeculativeDecoder:
    def __init__(self, target, draft, num_speculative_tokens: int = 4, max_sequence_length: int = 2048, vocab_size: Optional[int] = None):
        self.target, self.draft, self.k = target, draft, num_speculative_tokens
        self.max_len, self.vocab_size = max_sequence_length, vocab_size
        self.stats = SpeculativeStats()

    @staticmethod
    def _fwd(model, ctx: InferenceParams, tokens: List[int], start: int, dev) -> torch.Tensor:
        t = torch.tensor([tokens], device=dev)
        pos = torch.arange(start, start + len(tokens), device=dev)[None]
        ctx.sequence_len_offset = start
        out = model(t, pos, None, inference_context=ctx)[0]          # [n, v]
        ctx.sequence_len_offset = start + len(tokens)
        return out

    @torch.no_grad()
    def generate(self, prompt: List[int], params: Optional[SamplingParams] = None) -> List[int]:
        p = params or SamplingParams(temperature=0.0)
        self.target.eval()
        if hasattr(self.draft, "eval"):
            self.draft.eval()
        dev = next(self.target.parameters()).device
        gen = torch.Generator(device=dev)
        if p.seed is not None:
            gen.manual_seed(p.seed)
        tctx, dctx = InferenceParams(1, self.max_len), InferenceParams(1, self.max_len)
        seq = list(prompt)
        # prefill both caches on everything but the last prompt token; that token is the first input of the loop
        if len(seq) > 1:
            self._fwd(self.target, tctx, seq[:-1], 0, dev)
            self._fwd(self.draft, dctx, seq[:-1], 0, dev)
            self.stats.target_forwards += 1
            self.stats.draft_forwards += 1
        n_t = n_d = len(seq) - 1                                      # tokens resident in the target / draft cache
        out: List[int] = []
        while len(out) < p.num_tokens_to_generate:
            k = min(self.k, p.num_tokens_to_generate - len(out) - 1, self.max_len - len(seq) - 1)
            # ---- draft k tokens ----
            d_tokens, d_probs = [], []
            cur = seq[n_d:]                                           # tokens the draft cache has not seen yet (≥ 1)
            for _ in range(max(k, 0)):
                lg = self._fwd(self.draft, dctx, cur, n_d, dev)[-1]
                n_d += len(cur)
                self.stats.draft_forwards += 1
                pr = _probs(lg, p, self.vocab_size)
                tok = int(torch.multinomial(pr, 1, generator=gen)) if (pr > 0).sum() > 1 else int(pr.argmax())
                d_tokens.append(tok)
                d_probs.append(pr)
                cur = [tok]
            # ---- verify with one target forward over [unseen suffix, drafts] ----
            t_in = seq[n_t:] + d_tokens
            lg = self._fwd(self.target, tctx, t_in, n_t, dev)
            self.stats.target_forwards += 1
            t_probs = _probs(lg[len(t_in) - len(d_tokens) - 1 :], p, self.vocab_size)       # k + 1 rows: after the last real token and after each draft
            if d_tokens:
                n_acc, nxt = verify_draft_tokens(torch.tensor(d_tokens, device=dev), torch.stack(d_probs), t_probs, gen)
            else:
                n_acc, nxt = 0, (int(torch.multinomial(t_probs[0], 1, generator=gen)) if (t_probs[0] > 0).sum() > 1 else int(t_probs[0].argmax()))
            self.stats.proposed += len(d_tokens)
            self.stats.accepted += n_acc
            self.stats.per_step_accepted.append(n_acc)
            new = d_tokens[:n_acc] + [nxt]
            # ---- rewind: caches hold exactly the accepted history (the correction / bonus token is fed next round) ----
            n_t = len(seq) + n_acc
            n_d = min(n_d, len(seq) + n_acc)
            seq += new
            for t in new:
                out.append(t)
                if t in p.stop_token_ids or len(out) >= p.num_tokens_to_generate:
                    return out
        return out


class MTPDraft:
    """Use a model's multi-token-prediction heads as the draft: ``propose(hidden, last_token)`` chains the MTP layers to emit one token per head.
    Exposed with the plain model interface through ``__call__`` so ``SpeculativeDecoder`` can drive it; models without MTP heads fall back to
    their main head (acceptance 100 %, no speed-up — useful as a correctness oracle)."""

    def __init__(self, model):
        self.model = model

    def parameters(self):
        return self.model.parameters()

    def eval(self):
        self.model.eval()
        return self

    def __call__(self, tokens, position_ids, attention_mask, inference_context=None):
        return self.model(tokens, position_ids, attention_mask, inference_context=inference_context)
