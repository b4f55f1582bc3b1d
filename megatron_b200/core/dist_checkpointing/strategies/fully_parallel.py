"""Spread the writing of replicated shards over the ranks that hold them
(reference ``strategies/fully_parallel.py:46-165``).

Without this every DP-replicated tensor is written by the rank whose ``replica_id`` is all-zero
(DP rank 0), leaving the other DP ranks idle.  Here the ranks of ``parallelization_group``
exchange (key, offset, nbytes) of their shards and greedily assign each distinct shard to the
least-loaded holder; the chosen holder's replica id is rewritten to 0, everyone else's to ≠0."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

import torch
import torch.distributed as dist

from ..dict_utils import nested_values
from ..mapping import ShardedStateDict, ShardedTensor


def _shard_id(st: ShardedTensor):
    return (st.key, tuple(st.global_offset), tuple(st.local_shape))


def distribute_shards_to_ranks(shard_to_ranks: Dict, shard_to_size: Dict, num_ranks: int) -> Dict:
    """Greedy: rarest-holder-first, then largest-first, to the currently lightest holder."""
    load = [0] * num_ranks
    out = {}
    for sid in sorted(shard_to_ranks, key=lambda s: (len(shard_to_ranks[s]), -shard_to_size[s], str(s))):
        holders = shard_to_ranks[sid]
        r = min(holders, key=lambda x: (load[x], x))
        out[sid] = r
        load[r] += shard_to_size[sid]
    return out


class FullyParallelSaveStrategyWrapper:
    def __init__(self, strategy=None, parallelization_group=None, do_cache_distribution: bool = False):
        self.base_strategy = strategy
        self.group = parallelization_group
        self.do_cache_distribution = do_cache_distribution
        self._cached = None

    def apply_saving_parallelization(self, sharded_state_dict: ShardedStateDict) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            return
        ws = dist.get_world_size(self.group)
        if ws == 1:
            return
        rank = dist.get_rank(self.group)
        sts = [s for s in nested_values(sharded_state_dict) if isinstance(s, ShardedTensor)]
        # only shards that have a main replica *within this group's replication* are candidates:
        # we key on everything but the last replica-id component (the DP one)
        def base_replica(st):
            r = st.replica_id
            return r[:-1] if isinstance(r, tuple) else ()

        mine = [(_shard_id(s), base_replica(s), s.data.numel() * s.data.element_size() if s.data is not None else 0) for s in sts]
        gathered = [None] * ws
        dist.all_gather_object(gathered, mine, group=self.group)
        shard_to_ranks, shard_to_size = defaultdict(list), {}
        for r, lst in enumerate(gathered):
            for sid, base, nbytes in lst:
                if all(x == 0 for x in base):
                    shard_to_ranks[sid].append(r)
                    shard_to_size[sid] = nbytes
        assign = distribute_shards_to_ranks(shard_to_ranks, shard_to_size, ws)
        for s in sts:
            sid = _shard_id(s)
            if sid in assign and all(x == 0 for x in base_replica(s)):
                base = base_replica(s)
                s.replica_id = (*base, 0) if assign[sid] == rank else (*base, 1 + ((rank - assign[sid]) % ws))


class FullyParallelLoadStrategyWrapper:
    """Every shard that several ranks of ``parallelization_group`` request identically (DP replicas) is READ FROM STORAGE BY ONE of them and then exchanged
    over the group (reference ``strategies/fully_parallel.py:168-420`` + ``exchange_utils.py``): the checkpoint is read once per group instead of once per
    rank.  ``exchange_algo``: ``"broadcast"`` (one broadcast per shard from its reader) or ``"gather_object"`` (one all-gather of the readers' payloads,
    for many small shards)."""

    def __init__(self, strategy=None, parallelization_group=None, do_cache_distribution: bool = False, exchange_algo: str = "broadcast"):
        self.base_strategy, self.group = strategy, parallelization_group
        self.do_cache_distribution, self.exchange_algo = do_cache_distribution, exchange_algo
        self._cached = None
        self.last_stats: Dict[str, int] = {}

    def plan(self, sharded_tensors):
        """{shard id: reader rank} for shards requested by more than one rank of the group, balanced by bytes (cached if asked to)."""
        ws = dist.get_world_size(self.group)
        if self.do_cache_distribution and self._cached is not None:
            return self._cached
        mine = [(_shard_id(s), s.data.numel() * s.data.element_size() if s.data is not None else int(torch.tensor(s.local_shape).prod()) * 4) for s in sharded_tensors]
        gathered = [None] * ws
        dist.all_gather_object(gathered, mine, group=self.group)
        shard_to_ranks, shard_to_size = defaultdict(list), {}
        for r, lst in enumerate(gathered):
            for sid, nbytes in lst:
                shard_to_ranks[sid].append(r)
                shard_to_size[sid] = nbytes
        shared = {sid: rs for sid, rs in shard_to_ranks.items() if len(rs) > 1}
        assign = distribute_shards_to_ranks(shared, shard_to_size, ws)
        out = (assign, shared)
        if self.do_cache_distribution:
            self._cached = out
        return out

    def load(self, sharded_tensors, sharded_objects, checkpoint_dir: str, load_fn):
        """``load_fn(tensors, objects) -> objects`` reads from storage (fills ``st.data`` in place).  Returns the loaded objects."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return load_fn(sharded_tensors, sharded_objects)
        rank = dist.get_rank(self.group)
        for st in sharded_tensors:
            if st.data is None:
                st.init_data(device="cpu")
        assign, shared = self.plan(sharded_tensors)
        read_here = [s for s in sharded_tensors if _shard_id(s) not in assign or assign[_shard_id(s)] == rank]
        # several local ShardedTensors may name the same shard: read once, copy locally afterwards
        first_of: Dict = {}
        to_read = []
        for s in read_here:
            sid = _shard_id(s)
            if sid not in first_of:
                first_of[sid] = s
                to_read.append(s)
        objs = load_fn(to_read, sharded_objects)
        for s in read_here:
            src = first_of[_shard_id(s)]
            if s is not src:
                s.data.copy_(src.data)
        self.last_stats = {"read_bytes": sum(s.data.numel() * s.data.element_size() for s in to_read),
                           "received_bytes": 0, "shards_read": len(to_read), "shards_received": 0}
        # exchange, in a deterministic order known to every rank
        by_id = defaultdict(list)
        for s in sharded_tensors:
            by_id[_shard_id(s)].append(s)
        if self.exchange_algo == "gather_object":
            payload = {sid: first_of[sid].data for sid in assign if assign[sid] == rank and sid in first_of}
            gathered = [None] * dist.get_world_size(self.group)
            dist.all_gather_object(gathered, payload, group=self.group)
            for r, pl in enumerate(gathered):
                if r == rank:
                    continue
                for sid, t in pl.items():
                    for s in by_id.get(sid, []):
                        s.data.copy_(t)
                        self.last_stats["received_bytes"] += t.numel() * t.element_size()
                        self.last_stats["shards_received"] += 1
            return objs
        for sid in sorted(assign, key=str):
            reader = assign[sid]
            holders = shared[sid]
            if rank not in holders:
                continue
            mine = by_id[sid]
            buf = mine[0].data
            # broadcast inside the group; ranks that do not hold the shard skip it (they are not part of this shard's exchange), so use point-to-point
            # style broadcast over the holders only: the reader sends, holders receive
            if rank == reader:
                for h in holders:
                    if h != rank:
                        dist.send(buf.contiguous(), dst=dist.get_global_rank(self.group, h) if self.group is not None else h, group=self.group)
            else:
                tmp = torch.empty_like(buf, device="cpu") if buf.device.type != "cpu" and dist.get_backend(self.group) == "gloo" else buf
                dist.recv(tmp, src=dist.get_global_rank(self.group, reader) if self.group is not None else reader, group=self.group)
                for s in mine:
                    s.data.copy_(tmp)
                self.last_stats["received_bytes"] += buf.numel() * buf.element_size()
                self.last_stats["shards_received"] += 1
        return objs
