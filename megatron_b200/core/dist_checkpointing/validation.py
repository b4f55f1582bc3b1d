"""Integrity checks (reference ``dist_checkpointing/validation.py:45-456``)."""
from __future__ import annotations

from collections import defaultdict
from enum import Enum
from typing import Set

import numpy as np
import torch.distributed as dist

from .core import CheckpointingException
from .dict_utils import nested_values
from .mapping import ShardedBase, ShardedObject, ShardedStateDict, ShardedTensor, is_main_replica


class StrictHandling(Enum):
    ASSUME_OK_UNEXPECTED = "assume_ok_unexpected"
    LOG_UNEXPECTED = "log_unexpected"
    LOG_ALL = "log_all"
    RAISE_UNEXPECTED = "raise_unexpected"
    RAISE_ALL = "raise_all"
    RETURN_UNEXPECTED = "return_unexpected"
    RETURN_ALL = "return_all"
    IGNORE_ALL = "ignore_all"

    @staticmethod
    def requires_explicit_ckpt_mismatch_check(v: "StrictHandling") -> bool:
        return v != StrictHandling.ASSUME_OK_UNEXPECTED

    @staticmethod
    def requires_global_app_metadata(v: "StrictHandling") -> bool:
        return v in (StrictHandling.IGNORE_ALL, StrictHandling.RAISE_ALL, StrictHandling.RETURN_ALL, StrictHandling.LOG_ALL)

    @staticmethod
    def requires_returning_mismatch_keys(v: "StrictHandling") -> bool:
        return v in (StrictHandling.RETURN_UNEXPECTED, StrictHandling.RETURN_ALL)


def _shard_meta(sh: ShardedBase):
    if isinstance(sh, ShardedTensor):
        return ("T", sh.key, tuple(sh.global_shape), tuple(sh.global_offset), (1,) * sh.prepend_axis_num + tuple(sh.local_shape), sh.replica_id)
    if isinstance(sh, ShardedObject):
        return ("O", sh.unique_key, tuple(sh.global_shape), tuple(sh.global_offset), (), sh.replica_id)
    return None


def validate_sharding_integrity(sharded_state_dict: ShardedStateDict, process_group=None) -> None:
    """Every element of every global tensor must be written by exactly one main replica."""
    local = [m for m in (_shard_meta(s) for s in nested_values(sharded_state_dict)) if m is not None]
    if dist.is_available() and dist.is_initialized():
        ws = dist.get_world_size(process_group)
        gathered = [None] * ws
        dist.all_gather_object(gathered, local, group=process_group)
        if dist.get_rank(process_group) != 0:
            return
        allm = [m for g in gathered for m in g]
    else:
        allm = local
    by_key = defaultdict(list)
    for m in allm:
        by_key[m[1]].append(m)
    for key, shards in by_key.items():
        kind = shards[0][0]
        mains = [s for s in shards if is_main_replica(s[5])]
        if kind == "O":
            if len(mains) != 1:
                raise CheckpointingException(f"invalid access pattern for object {key}: {len(mains)} main replicas")
            continue
        gshape = shards[0][2]
        if any(s[2] != gshape for s in shards):
            raise CheckpointingException(f"global shape mismatch for {key}: {[s[2] for s in shards]}")
        total = int(np.prod(gshape)) if gshape else 1
        covered = 0
        seen = set()
        for s in mains:
            sig = (s[3], s[4])
            if sig in seen:
                raise CheckpointingException(f"shard {sig} of {key} is written by more than one main replica")
            seen.add(sig)
            covered += int(np.prod(s[4])) if s[4] else 1
        if covered != total:
            # allow_shape_mismatch tensors (padded vocab) may legitimately not tile: check overlap instead
            mains_sorted = sorted(mains, key=lambda s: s[3])
            if covered > total:
                raise CheckpointingException(f"invalid access pattern for {key}: {covered} elements written, global tensor has {total}")
            raise CheckpointingException(f"invalid access pattern for {key}: only {covered} of {total} elements are covered by main replicas")


def determine_global_metadata(sharded_state_dict):
    local = [m for m in (_shard_meta(s) for s in nested_values(sharded_state_dict)) if m is not None]
    if dist.is_available() and dist.is_initialized():
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, local)
        return local, out
    return local, [local]


def validate_integrity_and_strict_load(sharded_state_dict, strict: StrictHandling, ckpt_keys: Set[str]):
    """Compare requested keys with the checkpoint's; returns (missing_in_ckpt, unexpected_in_ckpt)."""
    requested = set()
    for s in nested_values(sharded_state_dict):
        if isinstance(s, ShardedTensor):
            requested.add(s.key)
        elif isinstance(s, ShardedObject):
            requested.add(s.unique_key)
    missing = requested - ckpt_keys
    unexpected = ckpt_keys - requested if StrictHandling.requires_global_app_metadata(strict) else set()
    if strict in (StrictHandling.RAISE_UNEXPECTED, StrictHandling.RAISE_ALL) and (missing or (strict == StrictHandling.RAISE_ALL and unexpected)):
        raise CheckpointingException(f"missing keys in checkpoint: {sorted(missing)[:10]}…; unexpected: {sorted(unexpected)[:10]}…")
    return missing, unexpected
