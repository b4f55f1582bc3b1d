"""Drop-in import compatibility: after ``megatron_b200.compat.install()``, ``import megatron.core…`` resolves
to ``megatron_b200.core…`` (and ``megatron.training`` → ``megatron_b200.training``), so scripts written against
the reference's public API (e.g. ``examples/run_simple_mcore_train_loop.py``) run unmodified on this framework."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

_PREFIXES = {"megatron.core": "megatron_b200.core", "megatron.training": "megatron_b200.training"}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "megatron":
            return None
        for alias, real in _PREFIXES.items():
            if fullname == alias or fullname.startswith(alias + "."):
                real_name = real + fullname[len(alias):]
                try:
                    if importlib.util.find_spec(real_name) is None:
                        return None
                except (ImportError, ValueError):
                    return None
                return importlib.util.spec_from_loader(fullname, _AliasLoader(real_name), is_package=True)
        return None


def install() -> None:
    if any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        return
    import types

    if "megatron" not in sys.modules:
        pkg = types.ModuleType("megatron")
        pkg.__path__ = []  # namespace-like package
        sys.modules["megatron"] = pkg
    sys.meta_path.insert(0, _AliasFinder())
    import megatron.core  # noqa: F401  (materialise the alias and attach it to the parent)

    sys.modules["megatron"].core = sys.modules["megatron.core"]
