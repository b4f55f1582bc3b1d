"""Rank helpers that work before ``torch.distributed`` is up (reference ``_rank_utils.py:15-93``): the launcher's environment
(torchrun, then SLURM) answers until the process group exists."""
from __future__ import annotations

import logging
import os

from ._slurm_utils import resolve_slurm_rank, resolve_slurm_world_size


def _dist():
    try:
        import torch.distributed as dist

        return dist if dist.is_available() and dist.is_initialized() else None
    except Exception:
        return None


def safe_get_rank() -> int:
    d = _dist()
    if d is not None:
        return d.get_rank()
    if "RANK" in os.environ:
        return int(os.environ["RANK"])
    r = resolve_slurm_rank()
    return 0 if r is None else r


def safe_get_world_size() -> int:
    d = _dist()
    if d is not None:
        return d.get_world_size()
    if "WORLD_SIZE" in os.environ:
        return int(os.environ["WORLD_SIZE"])
    w = resolve_slurm_world_size()
    return 1 if w is None else w


def log_single_rank(logger: logging.Logger, *args, rank: int = 0, **kwargs) -> None:
    """``logger.log(level, msg, ...)`` on one rank only (negative ranks count from the end: ``-1`` = last rank)."""
    if rank < 0:
        rank += safe_get_world_size()
    if safe_get_rank() == rank:
        logger.log(*args, **kwargs)
