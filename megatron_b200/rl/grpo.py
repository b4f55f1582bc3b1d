"""GRPO — group-relative policy optimisation (reference ``megatron/rl/`` 6.3 kLoC: agents, rollout server, GRPO loss in ``train_rl.py``).

Loop:  prompts → G sampled completions per prompt (the training model itself, through ``StaticInferenceEngine``) → scalar rewards from an
``Environment`` → advantages normalised inside each group → clipped-ratio policy-gradient loss with a KL penalty to the frozen reference
policy, evaluated with ONE teacher-forced forward of the training model over [prompt | completion].  The training model and the
rollout engine share weights in place (no refit/reshard step is needed on a single replica; across layouts use ``core.resharding``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from ..core.inference.engine import StaticInferenceEngine
from ..core.inference.sampling import SamplingParams


@dataclass
class GRPOConfig:
    group_size: int = 4
    max_new_tokens: int = 16
    temperature: float = 1.0
    clip_eps: float = 0.2
    kl_beta: float = 0.01
    entropy_coef: float = 0.0                      # weight of the (sampled-token) entropy bonus (reference --grpo-entropy-term-weight)
    clip_eps_upper: Optional[float] = None         # asymmetric clipping (reference --grpo-clamp-eps-upper); None = clip_eps
    filter_groups_with_same_reward: bool = False   # groups whose rollouts all scored the same carry no signal: drop them from the loss
    use_sequence_packing: bool = False             # log-probs / training on packed (THD) rows instead of padded ones (rl/sequence_packing_utils.py)
    packing_bin_size: int = 512


class Environment:
    """Reward model / verifier interface (reference ``rl/agent/api.py``)."""

    def prompts(self, n: int) -> List[List[int]]:
        raise NotImplementedError

    def reward(self, prompt: Sequence[int], completion: Sequence[int]) -> float:
        raise NotImplementedError


class CountTokenEnv(Environment):
    """Toy verifiable task: reward = fraction of completion tokens equal to ``target``."""

    def __init__(self, vocab: int, target: int = 7, prompt_len: int = 4, seed: int = 0):
        self.vocab, self.target, self.prompt_len = vocab, target, prompt_len
        self.g = torch.Generator().manual_seed(seed)

    def prompts(self, n):
        return [torch.randint(0, self.vocab, (self.prompt_len,), generator=self.g).tolist() for _ in range(n)]

    def reward(self, prompt, completion):
        return sum(1.0 for t in completion if t == self.target) / max(1, len(completion))


def sequence_logprobs(model, tokens: torch.Tensor, vocab_size: Optional[int] = None) -> torch.Tensor:
    """Teacher-forced log p(token_t | token_<t) for t ≥ 1;  tokens [b, s] → [b, s-1]."""
    b, s = tokens.shape
    pos = torch.arange(s, device=tokens.device)[None].expand(b, -1)
    logits = model(tokens, pos, None).float()          # [b, s, v]
    if vocab_size is not None:
        logits = logits[..., :vocab_size]
    lp = torch.log_softmax(logits[:, :-1], dim=-1)
    return lp.gather(-1, tokens[:, 1:].unsqueeze(-1)).squeeze(-1)


def group_advantages(rewards: torch.Tensor, group_size: int, eps: float = 1e-6) -> torch.Tensor:
    r = rewards.view(-1, group_size)
    return ((r - r.mean(dim=1, keepdim=True)) / (r.std(dim=1, keepdim=True) + eps)).view(-1)


def grpo_loss(logp, old_logp, ref_logp, advantages, mask, cfg: GRPOConfig):
    """All [b, s-1]; ``mask`` selects completion tokens.  Returns (loss, stats)."""
    ratio = torch.exp(logp - old_logp)
    adv = advantages.unsqueeze(-1)
    hi = cfg.clip_eps if cfg.clip_eps_upper is None else cfg.clip_eps_upper
    pg = -torch.min(ratio * adv, torch.clamp(ratio, 1 - cfg.clip_eps, 1 + hi) * adv)
    # unbiased low-variance KL estimator k3 = exp(ref - logp) - (ref - logp) - 1
    d = ref_logp - logp
    kl = torch.exp(d) - d - 1.0
    per_tok = pg + cfg.kl_beta * kl
    if cfg.entropy_coef:
        per_tok = per_tok + cfg.entropy_coef * logp          # minimising E[log p] of the sampled tokens = maximising the policy's entropy estimate
    denom = mask.sum().clamp(min=1)
    loss = (per_tok * mask).sum() / denom
    return loss, {"pg": ((pg * mask).sum() / denom).detach(), "kl": ((kl * mask).sum() / denom).detach(), "ratio_max": (ratio * mask).max().detach()}


class GRPOTrainer:
    def __init__(self, model, ref_model, optimizer, env: Environment, cfg: GRPOConfig, vocab_size: int, pad_id: int = 0):
        self.model, self.ref, self.opt, self.env, self.cfg, self.vocab, self.pad = model, ref_model, optimizer, env, cfg, vocab_size, pad_id
        for p in self.ref.parameters():
            p.requires_grad = False
        self.engine = StaticInferenceEngine(model, max_batch_size=256, max_sequence_length=512, vocab_size=vocab_size)
        from .rl_profiling import RLProfiler

        self.profiler = RLProfiler(enabled=False)

    @torch.no_grad()
    def rollout(self, n_prompts: int):
        prompts = self.env.prompts(n_prompts)
        rep = [p for p in prompts for _ in range(self.cfg.group_size)]
        self.model.eval()
        outs = self.engine.generate(rep, SamplingParams(temperature=self.cfg.temperature, num_tokens_to_generate=self.cfg.max_new_tokens))
        self.model.train()
        comps = [o[len(p):] for o, p in zip(outs, rep)]
        rewards = torch.tensor([self.env.reward(p, c) for p, c in zip(rep, comps)], dtype=torch.float32)
        L = max(len(o) for o in outs)
        dev = next(self.model.parameters()).device
        tokens = torch.full((len(outs), L), self.pad, dtype=torch.long, device=dev)
        mask = torch.zeros(len(outs), L - 1, device=dev)
        for i, (o, p) in enumerate(zip(outs, rep)):
            tokens[i, : len(o)] = torch.tensor(o, device=dev)
            mask[i, len(p) - 1 : len(o) - 1] = 1.0   # positions whose TARGET is a completion token
        return tokens, mask, rewards.to(dev)

    def _packed_logprobs(self, model, tokens: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """Same ``[b, s - 1]`` result as ``sequence_logprobs`` but computed on packed rows: padding never enters the model."""
        from .sequence_packing_utils import build_packed_batch, pack_sequences, packed_sequence_logprobs, unpack

        lens = [int(mask[i].nonzero().max()) + 2 if mask[i].any() else 1 for i in range(tokens.shape[0])]       # real length = last completion target + 1
        seqs = [tokens[i, :n].tolist() for i, n in enumerate(lens)]
        out = torch.zeros(tokens.shape[0], tokens.shape[1] - 1, device=tokens.device)
        for idx in pack_sequences(lens, self.cfg.packing_bin_size):
            pb = build_packed_batch(seqs, [1] * len(seqs), idx, device=tokens.device)
            for i, lp in zip(pb.seq_index, unpack(packed_sequence_logprobs(model, pb, self.vocab), pb)):
                out[i, : lp.numel()] = lp
        return out

    def step(self, n_prompts: int = 4, inner_epochs: int = 1):
        prof = self.profiler
        with prof.phase("rollout"):
            tokens, mask, rewards = self.rollout(n_prompts)
        prof.count("rollout", tokens=int(mask.sum()), sequences=tokens.shape[0])
        adv = group_advantages(rewards, self.cfg.group_size)
        if self.cfg.filter_groups_with_same_reward:
            r = rewards.view(-1, self.cfg.group_size)
            informative = (r.max(dim=1).values > r.min(dim=1).values).repeat_interleave(self.cfg.group_size)
            mask = mask * informative.unsqueeze(-1).to(mask.dtype)
        lp_fn = (lambda m: self._packed_logprobs(m, tokens, mask)) if self.cfg.use_sequence_packing else (lambda m: sequence_logprobs(m, tokens, self.vocab))
        with torch.no_grad(), prof.phase("logprobs", tokens=2 * tokens.numel()):
            old = lp_fn(self.model)
            ref = lp_fn(self.ref)
        stats = {}
        with prof.phase("train", tokens=inner_epochs * tokens.numel()):
            for _ in range(inner_epochs):
                logp = lp_fn(self.model)
                loss, stats = grpo_loss(logp, old, ref, adv, mask, self.cfg)
                self.opt.zero_grad()
                loss.backward()
                self.opt.step()
        stats.update(loss=loss.detach(), reward=rewards.mean())
        prof.end_iteration(reward=float(rewards.mean()))
        return stats
