"""Checkpoint-level metadata file ``metadata.json`` (reference ``dist_checkpointing/core.py:23-81``)."""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass
from typing import Optional

CONFIG_FNAME = "metadata.json"


class CheckpointingException(Exception):
    pass


@dataclass
class CheckpointingConfig:
    sharded_backend: str
    sharded_backend_version: int = 1
    common_backend: str = "torch"
    common_backend_version: int = 1


def check_is_distributed_checkpoint(checkpoint_dir) -> bool:
    return maybe_load_config(checkpoint_dir) is not None


def maybe_load_config(checkpoint_dir: str) -> Optional[CheckpointingConfig]:
    p = os.path.join(str(checkpoint_dir), CONFIG_FNAME)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return CheckpointingConfig(**json.load(f))


def save_config(config: CheckpointingConfig, checkpoint_dir: str):
    with open(os.path.join(str(checkpoint_dir), CONFIG_FNAME), "w") as f:
        json.dump(asdict(config), f)
