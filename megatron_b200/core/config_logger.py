"""Dump the configuration every module was built with (reference ``core/config_logger.py``): one JSON file per construction site under
``config.config_logger_dir`` so two runs can be diffed at the level of module kwargs instead of CLI flags."""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Any

import torch

_COUNTERS = {}


def has_config_logger_enabled(config) -> bool:
    return bool(getattr(config, "config_logger_dir", ""))


def get_config_logger_path(config) -> str:
    return getattr(config, "config_logger_dir", "")


def _jsonable(o: Any, depth: int = 0):
    if depth > 6:
        return repr(o)
    if o is None or isinstance(o, (bool, int, float, str)):
        return o
    if isinstance(o, torch.dtype):
        return str(o)
    if isinstance(o, torch.Tensor):
        return {"tensor": list(o.shape), "dtype": str(o.dtype)}
    if dataclasses.is_dataclass(o) and not isinstance(o, type):
        return {f.name: _jsonable(getattr(o, f.name), depth + 1) for f in dataclasses.fields(o)}
    if isinstance(o, dict):
        return {str(k): _jsonable(v, depth + 1) for k, v in o.items()}
    if isinstance(o, (list, tuple, set)):
        return [_jsonable(v, depth + 1) for v in o]
    if isinstance(o, torch.nn.Module):       # before the callable test: modules are callable
        return {"module": type(o).__qualname__, "children": {n: type(c).__qualname__ for n, c in o.named_children()}}
    if isinstance(o, type) or callable(o):
        return f"{getattr(o, '__module__', '')}.{getattr(o, '__qualname__', repr(o))}"
    return repr(o)


def log_config_to_disk(config, dict_data: dict, prefix: str = "", rank_str: str = "") -> str:
    path = get_config_logger_path(config)
    if not path:
        return ""
    os.makedirs(path, exist_ok=True)
    if not rank_str:
        rank_str = str(torch.distributed.get_rank()) if torch.distributed.is_available() and torch.distributed.is_initialized() else "0"
    dict_data = {k: v for k, v in dict_data.items() if k not in ("self", "__class__")}
    key = (path, prefix, rank_str)
    _COUNTERS[key] = _COUNTERS.get(key, 0) + 1
    fn = os.path.join(path, f"{prefix}.rank_{rank_str}.iter{_COUNTERS[key] - 1}.json")
    with open(fn, "w") as f:
        json.dump(_jsonable(dict_data), f, indent=1, sort_keys=True)
    return fn
