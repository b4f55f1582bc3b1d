"""SLURM environment probes (reference ``_slurm_utils.py:12-51``) for jobs started with ``srun`` instead of ``torchrun``."""
from __future__ import annotations

import os
from typing import Optional


def is_slurm_job() -> bool:
    return "SLURM_JOB_ID" in os.environ and "SLURM_PROCID" in os.environ


def _int(name: str) -> Optional[int]:
    v = os.environ.get(name)
    try:
        return int(v) if v is not None else None
    except ValueError:
        return None


def resolve_slurm_rank() -> Optional[int]:
    return _int("SLURM_PROCID") if is_slurm_job() else None


def resolve_slurm_world_size() -> Optional[int]:
    return _int("SLURM_NTASKS") if is_slurm_job() else None


def resolve_slurm_local_rank() -> Optional[int]:
    return _int("SLURM_LOCALID") if is_slurm_job() else None
