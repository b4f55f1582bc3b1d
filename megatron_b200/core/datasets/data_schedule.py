"""Sequence-packing data schedules over the data-parallel group (reference ``datasets/data_schedule.py:34-925`` + ``data_schedule_utils.py``).

Variable-length corpora (SFT, long-context mixes) make data-parallel ranks finish at very different times if each rank simply takes
the next ``micro_batch_size`` samples: the step ends when the rank that drew the longest documents is done.  The schedulers here look
at the lengths of ALL samples of a global batch (one small all-gather), decide which rank processes which sample, in which packed
micro-batch, move the samples there (one all-to-all of flat token tensors), and hand the training loop ready THD batches
(``tokens [1, T]`` + ``PackedSeqParams`` with ``cu_seqlens``).

* ``DpBalancedScheduler`` — longest-processing-time-first assignment of samples to DP ranks under the cost model
  ``len + len² / attention_scale`` (linear layers + causal attention), then first-fit-decreasing packing of each rank's samples into
  micro-batches of at most ``max_seqlen_per_rank`` tokens; every rank gets the same number of micro-batches (the pipeline schedule
  needs that), short ranks pad with an empty one.
* ``HybridCPDataLoaderWrapper`` — the same machinery with the hybrid DP×CP plan of ``pipeline_parallel/hybrid_cp_schedule.py``:
  samples longer than a rank's budget are given a CP sub-group instead of being split.

Everything that decides *where* a sample goes is computed identically on every rank from the gathered lengths — no coordinator.
"""
from __future__ import annotations

import enum
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..packed_seq_params import PackedSeqParams


def _pg_size(g) -> int:
    return dist.get_world_size(g) if dist.is_initialized() else 1


def _pg_rank(g) -> int:
    return dist.get_rank(g) if dist.is_initialized() else 0


class BasePackingScheduler:
    def __init__(self, max_seqlen_per_rank: int, dp_group=None, cp_size: int = 1, attention_scale: float = 4096.0, pad_to_multiple: int = 1):
        self.max_seqlen_per_rank, self.dp_group, self.cp_size = max_seqlen_per_rank, dp_group, cp_size
        self.attention_scale, self.pad_to_multiple = attention_scale, pad_to_multiple
        self.dp_size, self.dp_rank = _pg_size(dp_group), _pg_rank(dp_group)

    def get_required_sample_keys(self) -> Sequence[str]:
        return ("tokens",)

    def cost(self, n: int) -> float:
        return n + n * n / self.attention_scale

    def get_groups_and_subsamples(self, sample_id_seqlens: List[Tuple[int, int]]) -> List[List[List[int]]]:
        """``[(global sample id, length)]`` → ``plan[dp_rank][micro_batch] = [sample ids]``."""
        raise NotImplementedError

    # ---- exchange -------------------------------------------------------------------------
    def _gather_lengths(self, lens: List[int]) -> Tuple[List[Tuple[int, int]], List[int]]:
        """→ (``[(gid, len)]`` over all ranks, rank-major; first gid of every rank)."""
        if self.dp_size == 1:
            return [(i, n) for i, n in enumerate(lens)], [0, len(lens)]
        counts = [None] * self.dp_size
        dist.all_gather_object(counts, lens, group=self.dp_group)
        offs, out = [0], []
        for r, ls in enumerate(counts):
            out += [(offs[-1] + i, n) for i, n in enumerate(ls)]
            offs.append(offs[-1] + len(ls))
        return out, offs

    def _route(self, samples: List[Dict[str, torch.Tensor]], plan, offs, keys) -> Dict[int, Dict[str, torch.Tensor]]:
        """Move every sample to the rank the plan names; → ``{gid: sample}`` of what this rank now owns."""
        owners: Dict[int, List[int]] = {}
        for r, mbs in enumerate(plan):
            for mb in mbs:
                for gid in mb:
                    if r not in owners.setdefault(gid, []):
                        owners[gid].append(r)                       # a sample under context parallelism lives on several ranks
        mine0 = offs[self.dp_rank]
        if self.dp_size == 1:
            return {mine0 + i: s for i, s in enumerate(samples)}
        send: List[List[int]] = [[] for _ in range(self.dp_size)]
        for i in range(len(samples)):
            for r in owners.get(mine0 + i, []):
                send[r].append(i)
        meta_out = [[(mine0 + i, int(samples[i][keys[0]].numel())) for i in idxs] for idxs in send]
        meta_in = [None] * self.dp_size
        allm = [None] * self.dp_size                                 # (gid, numel) lists are tiny: gather all, keep our column
        dist.all_gather_object(allm, meta_out, group=self.dp_group)
        meta_in = [allm[r][self.dp_rank] for r in range(self.dp_size)]
        got: Dict[int, Dict[str, torch.Tensor]] = {gid: {} for m in meta_in for gid, _ in m}
        for k in keys:
            ref = samples[0][k] if samples else torch.zeros(0, dtype=torch.long)
            flat_out = torch.cat([samples[i][k].reshape(-1) for idxs in send for i in idxs]) if any(send) else ref.new_zeros(0)
            in_split = [sum(n for _, n in m) for m in meta_in]
            out_split = [sum(n for _, n in m) for m in meta_out]
            flat_in = flat_out.new_empty(sum(in_split))
            dist.all_to_all_single(flat_in, flat_out.contiguous(), in_split, out_split, group=self.dp_group)
            pos = 0
            for m in meta_in:
                for gid, n in m:
                    got[gid][k] = flat_in[pos : pos + n]
                    pos += n
        return got

    # ---- packing --------------------------------------------------------------------------
    def _pack(self, samples: List[Dict[str, torch.Tensor]], keys) -> Dict[str, Any]:
        lens = [int(s[keys[0]].numel()) for s in samples]
        m = self.pad_to_multiple
        padded = [-(-n // m) * m for n in lens]
        out: Dict[str, Any] = {}
        for k in keys:
            ref = samples[0][k] if samples else torch.zeros(0, dtype=torch.long)
            parts = []
            for s, n, p in zip(samples, lens, padded):
                parts.append(s[k].reshape(-1))
                if p > n:
                    parts.append(ref.new_zeros(p - n))
            out[k] = (torch.cat(parts) if parts else ref.new_zeros(0)).unsqueeze(0)
        cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
        cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0) if lens else cu[1:]
        cup = torch.zeros(len(lens) + 1, dtype=torch.int32)
        cup[1:] = torch.tensor(padded, dtype=torch.int32).cumsum(0) if lens else cup[1:]
        pos = torch.cat([torch.arange(p) for p in padded]) if padded else torch.zeros(0, dtype=torch.long)
        out["position_ids"] = pos.unsqueeze(0)
        valid = torch.cat([torch.arange(p) < n for n, p in zip(lens, padded)]) if padded else torch.zeros(0, dtype=torch.bool)
        if "loss_mask" in out:
            out["loss_mask"] = out["loss_mask"] * valid.to(out["loss_mask"].dtype).unsqueeze(0)
        out["padding_mask"] = (~valid).unsqueeze(0)
        # cu_seqlens of the UNPADDED sequences relative to the padded layout would not be contiguous: attention runs on the padded
        # boundaries, the padding rows are masked out of the loss
        out["packed_seq_params"] = PackedSeqParams(qkv_format="thd", cu_seqlens_q=cup, cu_seqlens_kv=cup, cu_seqlens_q_padded=cup, cu_seqlens_kv_padded=cup,
                                                   max_seqlen_q=max(padded, default=0), max_seqlen_kv=max(padded, default=0))
        out["seqlens"] = lens
        return out

    def run(self, local_samples: List[Dict[str, torch.Tensor]], keys: Optional[Sequence[str]] = None) -> List[Dict[str, Any]]:
        """One global batch: this rank's drawn samples in, this rank's packed micro-batches out."""
        keys = list(keys or (local_samples[0].keys() if local_samples else self.get_required_sample_keys()))
        lens = [int(s[keys[0]].numel()) for s in local_samples]
        all_lens, offs = self._gather_lengths(lens)
        plan = self.get_groups_and_subsamples(all_lens)
        got = self._route(local_samples, plan, offs, keys)
        return [self._pack([got[g] for g in mb], keys) for mb in plan[self.dp_rank]]


class DpBalancedScheduler(BasePackingScheduler):
    def get_groups_and_subsamples(self, sample_id_seqlens):
        budget = self.max_seqlen_per_rank * self.cp_size
        too_long = [(g, n) for g, n in sample_id_seqlens if n > budget]
        if too_long:
            raise ValueError(f"samples {too_long[:3]} exceed max_seqlen_per_rank × cp = {budget}; truncate them or use the hybrid-CP schedule")
        load = [0.0] * self.dp_size
        per_rank: List[List[Tuple[int, int]]] = [[] for _ in range(self.dp_size)]
        for g, n in sorted(sample_id_seqlens, key=lambda t: (-t[1], t[0])):       # LPT: longest first onto the least loaded rank
            r = min(range(self.dp_size), key=lambda i: (load[i], i))
            per_rank[r].append((g, n))
            load[r] += self.cost(n)
        plan: List[List[List[int]]] = []
        for items in per_rank:                                                     # first-fit decreasing into token-bounded micro-batches
            bins: List[Tuple[int, List[int]]] = []
            for g, n in items:
                for i, (used, ids) in enumerate(bins):
                    if used + n <= budget:
                        bins[i] = (used + n, ids + [g])
                        break
                else:
                    bins.append((n, [g]))
            plan.append([ids for _, ids in bins])
        n_mb = max(len(p) for p in plan)
        for p in plan:
            p += [[] for _ in range(n_mb - len(p))]
        return plan

    def imbalance(self, sample_id_seqlens, plan) -> float:
        """max / mean modelled cost over ranks (1.0 = perfect)."""
        n = dict(sample_id_seqlens)
        loads = [sum(self.cost(n[g]) for mb in p for g in mb) for p in plan]
        return max(loads) / (sum(loads) / len(loads)) if sum(loads) else 1.0


class NaiveSequentialScheduler(BasePackingScheduler):
    """Baseline: samples stay on the rank that drew them, packed in arrival order."""

    def get_groups_and_subsamples(self, sample_id_seqlens):
        budget = self.max_seqlen_per_rank * self.cp_size
        per = max(1, len(sample_id_seqlens) // self.dp_size)
        plan = []
        for r in range(self.dp_size):
            mine = sample_id_seqlens[r * per : (r + 1) * per] if r < self.dp_size - 1 else sample_id_seqlens[r * per :]
            bins, used = [[]], 0
            for g, n in mine:
                if used + n > budget and bins[-1]:
                    bins.append([])
                    used = 0
                bins[-1].append(g)
                used += n
            plan.append(bins)
        n_mb = max(len(p) for p in plan)
        for p in plan:
            p += [[] for _ in range(n_mb - len(p))]
        return plan


class PackingSchedulerEnum(enum.Enum):
    DP_BALANCED = "dp_balanced"
    NAIVE_SEQUENTIAL = "naive_sequential"


_SCHEDULERS = {PackingSchedulerEnum.DP_BALANCED: DpBalancedScheduler, PackingSchedulerEnum.NAIVE_SEQUENTIAL: NaiveSequentialScheduler}


class PackedBatchIterator:
    """Wraps a per-sample iterator: draws ``samples_per_rank`` samples per global batch, runs the scheduler, yields packed micro-batches.
    ``num_microbatches`` of the batch just scheduled tells the training loop how many to run this step."""

    def __init__(self, data_iterator: Iterator[Dict[str, torch.Tensor]], scheduler: BasePackingScheduler, samples_per_rank: int, keys: Optional[Sequence[str]] = None):
        self.it, self.sched, self.n, self.keys = data_iterator, scheduler, samples_per_rank, keys
        self._queue: List[Dict[str, Any]] = []
        self.num_microbatches = 0

    def __iter__(self):
        return self

    def __next__(self):
        if not self._queue:
            local = [next(self.it) for _ in range(self.n)]
            local = [{k: v for k, v in s.items() if torch.is_tensor(v) and v.dim() <= 1} for s in local]
            self._queue = self.sched.run(local, self.keys)
            self.num_microbatches = len(self._queue)
        return self._queue.pop(0)


def wrap_data_iterator(data_iterator, scheduler_type=PackingSchedulerEnum.DP_BALANCED, *, max_seqlen_per_rank: int, samples_per_rank: int, dp_group=None, cp_size: int = 1,
                       pad_to_multiple: int = 1, keys: Optional[Sequence[str]] = None) -> PackedBatchIterator:
    if isinstance(scheduler_type, str):
        scheduler_type = PackingSchedulerEnum(scheduler_type)
    sched = _SCHEDULERS[scheduler_type](max_seqlen_per_rank, dp_group, cp_size, pad_to_multiple=pad_to_multiple)
    return PackedBatchIterator(data_iterator, sched, samples_per_rank, keys)


def get_batch_on_this_rank_for_sequence_packing(packed_iterator: PackedBatchIterator, device=None) -> Dict[str, Any]:
    """The ``get_batch`` of a packed run: next micro-batch, tensors on ``device``, ``cu_seqlens`` too."""
    b = next(packed_iterator)
    if device is not None:
        for k, v in list(b.items()):
            if torch.is_tensor(v):
                b[k] = v.to(device, non_blocking=True)
        p = b["packed_seq_params"]
        for f in ("cu_seqlens_q", "cu_seqlens_kv", "cu_seqlens_q_padded", "cu_seqlens_kv_padded"):
            setattr(p, f, getattr(p, f).to(device, non_blocking=True))
    return b


class HybridCPDataLoaderWrapper:
    """Hybrid DP×CP flavour (reference :76-353): the plan comes from ``BalancedCPScheduler``, which gives a sample that exceeds one
    rank's token budget a CP sub-group of 2ᵏ ranks instead of truncating it.  Samples are routed with the same all-to-all — one
    copy to every rank of the sample's CP block — and each packed micro-batch carries, per sample, the ranks of its CP block
    (``cp_ranks``) so attention picks the matching process group (``PackedSeqParams.local_cp_size``)."""

    def __init__(self, data_iterator, max_seqlen_per_rank: int, samples_per_rank: int, dp_cp_group=None, keys: Optional[Sequence[str]] = None):
        from ..pipeline_parallel.hybrid_cp_schedule import BalancedCPScheduler

        self.it, self.n, self.keys = data_iterator, samples_per_rank, keys
        self._base = BasePackingScheduler(max_seqlen_per_rank, dp_cp_group)
        self.sched = BalancedCPScheduler(max_seqlen_per_rank, total_gpus=self._base.dp_size)
        self._queue: List[Dict[str, Any]] = []
        self.num_microbatches = 0

    def __iter__(self):
        return self

    def __next__(self):
        if not self._queue:
            local = [next(self.it) for _ in range(self.n)]
            local = [{k: v for k, v in s.items() if torch.is_tensor(v) and v.dim() <= 1} for s in local]
            keys = list(self.keys or local[0].keys())
            b = self._base
            all_lens, offs = b._gather_lengths([int(s[keys[0]].numel()) for s in local])
            groups = self.sched.get_groups_and_subsamples(all_lens)
            plan = [[list(g.per_gpu[r]) for g in groups] for r in range(b.dp_size)]
            got = b._route(local, plan, offs, keys)
            for g in groups:
                mine = list(g.per_gpu[b.dp_rank])
                batch = b._pack([got[i] for i in mine], keys)
                batch["sample_ids"] = mine
                batch["cp_ranks"] = [g.cp_ranks(i) for i in mine]
                sizes = {len(r) for r in batch["cp_ranks"]}
                batch["packed_seq_params"].local_cp_size = sizes.pop() if len(sizes) == 1 else None
                self._queue.append(batch)
            self.num_microbatches = len(self._queue)
        return self._queue.pop(0)
