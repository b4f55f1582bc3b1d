"""Module specs: describe *which class* implements each sub-module.

Same contract as reference ``transformer/spec_utils.py:10,76`` (``ModuleSpec`` with
``module``/``params``/``submodules`` and ``build_module``).
"""
from __future__ import annotations

import importlib
import types
from dataclasses import dataclass, field
from typing import Any, Tuple, Union


@dataclass
class ModuleSpec:
    module: Union[Tuple[str, str], type, Any]
    params: dict = field(default_factory=dict)
    submodules: Any = None
    metainfo: dict = field(default_factory=dict)


def import_module(path: Tuple[str, str]):
    mod_path, name = path
    mod = importlib.import_module(mod_path)
    return getattr(mod, name)


def get_module(spec_or_module, **extra):
    if isinstance(spec_or_module, ModuleSpec):
        m = spec_or_module.module
        return import_module(m) if isinstance(m, tuple) else m
    return spec_or_module


def build_module(spec_or_module, *args, **kwargs):
    """Instantiate whatever ``spec_or_module`` designates.

    Accepts a plain callable (returned as-is when it is a function), a class,
    or a ``ModuleSpec`` whose ``params`` are merged under the call kwargs and
    whose ``submodules`` are forwarded as ``submodules=``.
    """
    if isinstance(spec_or_module, types.FunctionType):
        return spec_or_module
    if isinstance(spec_or_module, ModuleSpec) and isinstance(spec_or_module.module, types.FunctionType):
        return spec_or_module.module
    if isinstance(spec_or_module, type):
        module, params, submodules = spec_or_module, {}, None
    elif isinstance(spec_or_module, ModuleSpec):
        module = get_module(spec_or_module)
        params = dict(spec_or_module.params or {})
        submodules = spec_or_module.submodules
    else:
        raise TypeError(f"cannot build module from {type(spec_or_module)}")
    if isinstance(module, types.FunctionType):
        return module
    if submodules is not None:
        kwargs = dict(kwargs, submodules=submodules)
    try:
        return module(*args, **{**params, **kwargs})
    except Exception as e:
        import sys

        raise type(e)(f"{e} when instantiating {getattr(module, '__name__', module)}").with_traceback(sys.exc_info()[2])
