"""Agents and the rollout bank (reference ``megatron/rl/agent/`` — ``Agent``, ``RewardOnlyAgent``, grouped rollouts — and ``megatron/rl/rl_utils.py`` rollout
collection with staleness control).

An *agent* turns prompts into scored trajectories by calling an ``InferenceInterface``; ``RolloutBank`` decouples generation from training: generation
fills it (possibly several policy versions behind), the trainer samples complete GROUPS (GRPO needs all samples of a prompt together) whose policy lag is
within ``max_staleness`` and drops the rest."""
from __future__ import annotations

import random
from abc import ABC, abstractmethod
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional, Sequence

from ..core.inference.sampling import SamplingParams
from .inference_interface import InferenceInterface, InferenceRequest


@dataclass
class Rollout:
    prompt: List[int]
    completion: List[int]
    reward: float
    policy_version: int = 0
    group_id: int = 0
    info: Dict = field(default_factory=dict)


class Agent(ABC):
    """``get_prompts`` + ``score``; ``rollout_groups`` does the generation round trip."""

    @abstractmethod
    def get_prompts(self, n: int) -> List[List[int]]:
        ...

    @abstractmethod
    def score(self, prompt: Sequence[int], completion: Sequence[int]) -> float:
        ...

    def rollout_groups(self, inference: InferenceInterface, n_prompts: int, group_size: int, sampling: SamplingParams, first_group_id: int = 0) -> List[List[Rollout]]:
        prompts = self.get_prompts(n_prompts)
        resps = inference.generate([InferenceRequest(p, sampling, n=group_size) for p in prompts])
        groups = []
        for gi, r in enumerate(resps):
            comps = [c[len(r.prompt_tokens):] if c[: len(r.prompt_tokens)] == list(r.prompt_tokens) else c for c in r.completions]
            groups.append([Rollout(list(r.prompt_tokens), list(c), float(self.score(r.prompt_tokens, c)), r.policy_version, first_group_id + gi) for c in comps])
        return groups


class RewardOnlyAgent(Agent):
    """An agent defined by a prompt sampler and a reward function (reference ``RewardOnlyAgent``)."""

    def __init__(self, prompt_fn, reward_fn):
        self.prompt_fn, self.reward_fn = prompt_fn, reward_fn

    def get_prompts(self, n):
        return self.prompt_fn(n)

    def score(self, prompt, completion):
        return self.reward_fn(prompt, completion)


class WeightedMultiAgent(Agent):
    """Mixture of agents (task curriculum): prompts are drawn from the sub-agents in proportion to their weights; the scorer is the prompt's owner."""

    def __init__(self, agents: Sequence[Agent], weights: Optional[Sequence[float]] = None, seed: int = 0):
        self.agents = list(agents)
        w = list(weights) if weights is not None else [1.0] * len(self.agents)
        self.weights = [x / sum(w) for x in w]
        self.rng = random.Random(seed)
        self._owner: Dict[tuple, Agent] = {}

    def get_prompts(self, n):
        out = []
        for _ in range(n):
            a = self.rng.choices(self.agents, self.weights)[0]
            p = a.get_prompts(1)[0]
            self._owner[tuple(p)] = a
            out.append(p)
        return out

    def score(self, prompt, completion):
        return self._owner[tuple(prompt)].score(prompt, completion)


class RolloutBank:
    def __init__(self, capacity_groups: int = 1024, max_staleness: int = 1, seed: int = 0):
        self.groups: Deque[List[Rollout]] = deque(maxlen=capacity_groups)
        self.max_staleness = max_staleness
        self.rng = random.Random(seed)
        self.dropped_stale = 0
        self._next_group = 0

    def next_group_id(self, n: int) -> int:
        g = self._next_group
        self._next_group += n
        return g

    def add(self, groups: List[List[Rollout]]) -> None:
        self.groups.extend(g for g in groups if g)

    def __len__(self) -> int:
        return len(self.groups)

    def sample(self, n_groups: int, current_version: int, drop_uninformative: bool = True) -> List[List[Rollout]]:
        """Take up to ``n_groups`` fresh-enough groups (oldest first, so nothing starves), removing them from the bank.  Groups whose rewards are all equal carry no
        GRPO signal (zero advantages) and are skipped when ``drop_uninformative``."""
        out, keep = [], deque()
        while self.groups:
            g = self.groups.popleft()
            if current_version - min(r.policy_version for r in g) > self.max_staleness:
                self.dropped_stale += 1
                continue
            if len(out) < n_groups and not (drop_uninformative and len({r.reward for r in g}) == 1):
                out.append(g)
            else:
                keep.append(g)
        self.groups.extend(keep)
        return out
