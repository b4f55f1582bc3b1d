"""Hybrid (per-sample) context parallelism: pack variable-length samples onto a DP×CP pool of GPUs
(reference ``pipeline_parallel/hybrid_cp_schedule.py`` — ``BalancedCPScheduler`` :14, ``hybrid_context_parallel_forward_backward`` :477).

Every sample gets the smallest power-of-two CP size that fits the per-rank token budget; a *group* is a set of samples
executed together by the whole pool (one forward-backward), and groups run back to back.  The packer here is a
target-driven longest-processing-time heuristic:

1. samples are ordered by per-GPU work (``len² / cp``), largest first;
2. the first sample of a group fixes the target ``T`` (nobody can finish earlier than the largest indivisible piece);
3. each further sample goes to the least-loaded *aligned* block of ``cp`` GPUs that has the token budget for it, unless that
   would push the block past ``T·(1+slack)`` — then it waits for the next group;
4. when GPUs are left idle, the heaviest samples are *widened* (CP size doubled onto an empty buddy block), which lowers
   the group's critical path instead of leaving silicon dark.

The result is a pure function of the sequence lengths, so every rank computes the same plan without communication.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import ceil, log2
from typing import Dict, Iterator, List, Optional, Sequence, Tuple


@dataclass
class HybridCPGroup:
    per_gpu: List[List[int]]                                        # sample ids resident on each GPU of the pool
    placement: Dict[int, Tuple[int, int]] = field(default_factory=dict)   # sample id -> (first GPU, CP size)

    def cp_ranks(self, sample_id: int) -> List[int]:
        start, size = self.placement[sample_id]
        return list(range(start, start + size))


class BalancedCPScheduler:
    def __init__(self, max_seq_len_per_rank: int, total_gpus: Optional[int] = None, dp_cp_group=None, slack: float = 0.15):
        if total_gpus is None:
            total_gpus = dp_cp_group.size()
        if total_gpus & (total_gpus - 1):
            raise ValueError("the DPxCP pool must be a power of two so CP blocks can be aligned")
        self.max_seq_len_per_rank, self.total_gpus, self.slack = max_seq_len_per_rank, total_gpus, slack

    # ---- cost model ----
    def gpus_needed(self, seq_len: int) -> int:
        n = max(1, 2 ** ceil(log2(max(seq_len / self.max_seq_len_per_rank, 1e-9)))) if seq_len > self.max_seq_len_per_rank else 1
        if n > self.total_gpus:
            raise ValueError(f"a sample of {seq_len} tokens needs {n} GPUs at {self.max_seq_len_per_rank} tokens/rank; the pool has {self.total_gpus}")
        return n

    @staticmethod
    def workload(seq_len: int, cp_size: int) -> float:
        """Relative per-GPU cost: attention dominates at the lengths where CP is used."""
        return seq_len * seq_len / cp_size

    def get_total_workload(self, seq_len: int, cp_size: Optional[int] = None) -> float:
        return self.workload(seq_len, cp_size or self.gpus_needed(seq_len))

    # ---- packing ----
    def _next_group(self, pending: List[Tuple[int, int]]) -> Tuple[HybridCPGroup, List[Tuple[int, int]]]:
        G = self.total_gpus
        load = [0.0] * G
        tokens = [0.0] * G
        grp = HybridCPGroup(per_gpu=[[] for _ in range(G)])
        lens: Dict[int, int] = {}
        left: List[Tuple[int, int]] = []
        target = None
        for sid, n in pending:
            cp = self.gpus_needed(n)
            w, tok = self.workload(n, cp), n / cp
            best, best_load = None, None
            for start in range(0, G, cp):
                blk = range(start, start + cp)
                if any(tokens[r] + tok > self.max_seq_len_per_rank + 1e-9 for r in blk):
                    continue
                m = max(load[r] for r in blk)
                if best is None or m < best_load:
                    best, best_load = start, m
            if best is None or (target is not None and best_load > 0 and best_load + w > target * (1 + self.slack)):
                left.append((sid, n))
                continue
            if target is None:
                target = w
            for r in range(best, best + cp):
                load[r] += w
                tokens[r] += tok
                grp.per_gpu[r].append(sid)
            grp.placement[sid] = (best, cp)
            lens[sid] = n
        self._widen(grp, lens, load, tokens)
        return grp, left

    def _widen(self, grp: HybridCPGroup, lens: Dict[int, int], load: List[float], tokens: List[float]) -> None:
        """Double the CP size of the heaviest samples onto empty buddy blocks while any exist."""
        changed = True
        while changed:
            changed = False
            for sid in sorted(grp.placement, key=lambda s: -self.workload(lens[s], grp.placement[s][1])):
                start, cp = grp.placement[sid]
                if cp * 2 > self.total_gpus:
                    continue
                buddy = start ^ cp                      # the other half of the aligned 2·cp block
                if any(grp.per_gpu[r] for r in range(buddy, buddy + cp)):
                    continue
                if any(len(grp.per_gpu[r]) != 1 for r in range(start, start + cp)):
                    continue                            # only samples that own their block are widened
                n = lens[sid]
                new_start, new_cp = min(start, buddy), cp * 2
                for r in range(new_start, new_start + new_cp):
                    grp.per_gpu[r] = [sid]
                    load[r] = self.workload(n, new_cp)
                    tokens[r] = n / new_cp
                grp.placement[sid] = (new_start, new_cp)
                changed = True
                break

    def get_groups_and_subsamples(self, sample_id_seqlens: Sequence[Tuple[int, int]], config=None) -> List[HybridCPGroup]:
        pending = sorted(sample_id_seqlens, key=lambda x: (-self.workload(x[1], self.gpus_needed(x[1])), x[0]))
        groups = []
        while pending:
            grp, pending = self._next_group(pending)
            groups.append(grp)
        return groups


def hybrid_context_parallel_forward_backward(*, forward_backward_one_sample, samples: Sequence[dict], scheduler: BalancedCPScheduler, rank_in_pool: int,
                                             cp_group_for=None) -> Iterator:
    """Drive one global batch: for each group, run this rank's samples with the CP group chosen for them.

    ``forward_backward_one_sample(sample, cp_ranks, cp_group)`` does the work; ``cp_group_for(cp_ranks)`` maps the rank list
    to a process group (``parallel_state.get_hybrid_context_parallel_group`` keeps one per power-of-two block).  Yields the
    per-sample results in execution order."""
    plan = scheduler.get_groups_and_subsamples([(i, int(s["seq_len"])) for i, s in enumerate(samples)])
    for grp in plan:
        for sid in grp.per_gpu[rank_in_pool]:
            ranks = grp.cp_ranks(sid)
            yield forward_backward_one_sample(samples[sid], ranks, cp_group_for(ranks) if cp_group_for is not None else None)
