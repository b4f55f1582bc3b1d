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


# ---- round-2 additions: strict-flag parsing, non-strict load adjustment, mismatch reports, file integrity manifest -------
import hashlib as _hashlib
import json as _json
import logging as _logging
import os as _os
from concurrent.futures import ThreadPoolExecutor as _Pool

_log = _logging.getLogger(__name__)
INTEGRITY_FNAME = "integrity.json"          # same file name and schema as the reference, so either side can verify the other's
_HASH_ALGORITHM = "sha256"


def parse_strict_flag(strict) -> StrictHandling:
    """``'log_all'`` / ``StrictHandling.LOG_ALL`` -> the enum (reference ``validation.py:107``)."""
    if isinstance(strict, StrictHandling):
        return strict
    try:
        return StrictHandling(strict)
    except ValueError as e:
        raise ValueError(f"invalid strict flag '{strict}': one of {[s.value for s in StrictHandling]}") from e


def verify_checkpoint(checkpoint_dir: str) -> None:
    from .core import check_is_distributed_checkpoint
    if not _os.path.isdir(str(checkpoint_dir)):
        raise CheckpointingException(f"Checkpoint directory {checkpoint_dir} does not exist")
    if not check_is_distributed_checkpoint(checkpoint_dir):
        raise CheckpointingException(f"{checkpoint_dir} is not a distributed checkpoint")


def adjust_non_strict_load(sharded_state_dict: ShardedStateDict, sharded_keys_to_remove: Set[str]) -> ShardedStateDict:
    """Drop the requested entries the checkpoint does not have, so that a non-strict load leaves those tensors untouched."""
    def keep(x):
        if isinstance(x, ShardedTensor):
            return x.key not in sharded_keys_to_remove
        if isinstance(x, ShardedObject):
            return x.unique_key not in sharded_keys_to_remove and x.key not in sharded_keys_to_remove
        return True

    def walk(d):
        if isinstance(d, dict):
            return {k: walk(v) for k, v in d.items() if keep(v)}
        if isinstance(d, list):
            return [walk(v) for v in d if keep(v)]
        return d
    return walk(sharded_state_dict)


def maybe_report_missing_and_unexpected_keys(strict: StrictHandling, missing_keys: Set[str], unexpected_keys: Set[str], raise_error: bool = True) -> None:
    """Log or raise according to the strictness (reference ``validation.py:285``).  "missing" = requested but not in the
    checkpoint is reported under the *unexpected* policies (the application expected them), as in the reference."""
    if not missing_keys and not unexpected_keys:
        return
    parts = []
    if missing_keys:
        parts.append(f"Missing keys (in the checkpoint, not requested): {sorted(missing_keys)[:20]}{' …' if len(missing_keys) > 20 else ''}")
    if unexpected_keys:
        parts.append(f"Unexpected keys (requested, not in the checkpoint): {sorted(unexpected_keys)[:20]}{' …' if len(unexpected_keys) > 20 else ''}")
    msg = "Some keys found in the checkpoint are missing in the provided sharded state dict or vice versa. " + " ".join(parts)
    raise_all = strict == StrictHandling.RAISE_ALL and (missing_keys or unexpected_keys)
    raise_unexp = strict == StrictHandling.RAISE_UNEXPECTED and unexpected_keys
    if raise_error and (raise_all or raise_unexp):
        raise CheckpointingException(msg)
    if strict in (StrictHandling.LOG_ALL, StrictHandling.LOG_UNEXPECTED, StrictHandling.RAISE_ALL, StrictHandling.RAISE_UNEXPECTED):
        _log.warning(msg)


def _compute_file_hash(path: str, chunk: int = 8 << 20) -> str:
    h = _hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def _manifest_files(checkpoint_dir: str):
    return sorted(e for e in _os.listdir(checkpoint_dir) if e != INTEGRITY_FNAME and _os.path.isfile(_os.path.join(checkpoint_dir, e)))


def save_integrity_manifest(checkpoint_dir: str, workers: int = 8) -> None:
    """SHA-256 of every file of the checkpoint -> ``integrity.json`` (call on ONE rank after the save has been finalized).
    Hashing releases the GIL, so a small thread pool reads the shard files concurrently (NVMe / parallel FS bound)."""
    names = _manifest_files(checkpoint_dir)
    with _Pool(max(1, min(workers, len(names) or 1))) as pool:
        digests = list(pool.map(lambda n: _compute_file_hash(_os.path.join(checkpoint_dir, n)), names))
    tmp = _os.path.join(checkpoint_dir, INTEGRITY_FNAME + ".tmp")
    with open(tmp, "w") as f:
        _json.dump({"algorithm": _HASH_ALGORITHM, "files": dict(zip(names, digests))}, f, indent=2)
    _os.replace(tmp, _os.path.join(checkpoint_dir, INTEGRITY_FNAME))


def verify_integrity_manifest(checkpoint_dir: str, workers: int = 8) -> None:
    """Raise ``CheckpointingException`` listing every missing, extra or corrupted file.  A checkpoint without a manifest
    passes with a warning (older checkpoints)."""
    path = _os.path.join(checkpoint_dir, INTEGRITY_FNAME)
    if not _os.path.isfile(path):
        _log.warning("no %s in %s: integrity not verified", INTEGRITY_FNAME, checkpoint_dir)
        return
    with open(path) as f:
        payload = _json.load(f)
    if payload.get("algorithm") != _HASH_ALGORITHM:
        raise CheckpointingException(f"integrity manifest uses unsupported algorithm {payload.get('algorithm')!r}")
    want = payload.get("files", {})
    have = set(_manifest_files(checkpoint_dir))
    problems = [f"missing file: {n}" for n in sorted(set(want) - have)] + [f"file not in manifest: {n}" for n in sorted(have - set(want))]
    names = sorted(set(want) & have)
    with _Pool(max(1, min(workers, len(names) or 1))) as pool:
        for n, d in zip(names, pool.map(lambda n: _compute_file_hash(_os.path.join(checkpoint_dir, n)), names)):
            if d != want[n]:
                problems.append(f"hash mismatch: {n}")
    if problems:
        raise CheckpointingException(f"checkpoint {checkpoint_dir} failed the integrity check: " + "; ".join(problems))
