"""Spread the writing of replicated shards over the ranks that hold them
(reference ``strategies/fully_parallel.py:46-165``).

Without this every DP-replicated tensor is written by the rank whose ``replica_id`` is all-zero
(DP rank 0), leaving the other DP ranks idle.  Here the ranks of ``parallelization_group``
exchange (key, offset, nbytes) of their shards and greedily assign each distinct shard to the
least-loaded holder; the chosen holder's replica id is rewritten to 0, everyone else's to ≠0."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

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
    """Load-side counterpart: API parity; DCP already reads each needed byte range once per rank."""

    def __init__(self, strategy=None, parallelization_group=None, do_cache_distribution: bool = False, exchange_algo: str = "broadcast"):
        self.base_strategy, self.group = strategy, parallelization_group
