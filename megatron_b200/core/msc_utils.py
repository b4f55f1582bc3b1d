"""Optional multi-storage-client backend (reference ``msc_utils.py:16-111``).

``MultiStorageClientFeature`` is a process-wide switch; when on (and the ``multistorageclient`` package is importable) checkpoint and
dataset code opens ``msc://profile/path`` URLs through it, otherwise the same attribute names resolve to their ``os`` / ``open`` /
``torch`` equivalents so call sites need one code path: ``msc = MultiStorageClientFeature.import_package(); msc.open(path)``."""
from __future__ import annotations

import glob as _glob
import os
import shutil
from typing import Any

import torch


class _LocalBackend:
    """The subset of the multistorageclient module API our code uses, on the local file system."""

    open = staticmethod(open)
    glob = staticmethod(_glob.glob)

    class os:                                   # noqa: N801 - mirrors ``msc.os``
        path = os.path
        makedirs = staticmethod(os.makedirs)
        listdir = staticmethod(os.listdir)
        remove = staticmethod(os.remove)
        rename = staticmethod(os.replace)

    class torch:                                # noqa: N801 - mirrors ``msc.torch``
        load = staticmethod(torch.load)
        save = staticmethod(torch.save)

    class numpy:                                # noqa: N801
        @staticmethod
        def load(path, **kw):
            import numpy

            return numpy.load(path, **kw)

        @staticmethod
        def memmap(path, **kw):
            import numpy

            return numpy.memmap(path, **kw)

    @staticmethod
    def Path(p):                                # noqa: N802
        import pathlib

        return pathlib.Path(p)

    @staticmethod
    def delete(path, recursive: bool = False):
        shutil.rmtree(path) if recursive and os.path.isdir(path) else os.remove(path)


class _FeatureFlag:
    def __init__(self, default: bool = False):
        self._enabled = default

    def enable(self) -> None:
        self._enabled = True

    def disable(self) -> None:
        self._enabled = False

    def is_enabled(self) -> bool:
        return self._enabled

    def import_package(self) -> Any:
        """The real package when enabled and installed; the local stand-in otherwise (never raises at call sites that only touch local paths)."""
        if self._enabled:
            try:
                import multistorageclient as msc

                return msc
            except ImportError as e:
                raise ImportError("MultiStorageClientFeature is enabled but the multistorageclient package is not installed") from e
        return _LocalBackend

    def is_msc_url(self, path) -> bool:
        return isinstance(path, str) and path.startswith("msc://")

    def __getstate__(self):
        return {"enabled": self._enabled}

    def __setstate__(self, state):
        self._enabled = state["enabled"]


MultiStorageClientFeature = _FeatureFlag(default=False)


class MaybeMultiStorageClient:
    """Attribute proxy: ``MaybeMultiStorageClient().open`` is ``msc.open`` or ``open`` depending on the switch, decided per call."""

    def path_isdir(self, path, strict: bool = True) -> bool:
        return MultiStorageClientFeature.import_package().os.path.isdir(path)

    def __getattr__(self, name):
        return getattr(MultiStorageClientFeature.import_package(), name)

    def __dir__(self):
        return dir(MultiStorageClientFeature.import_package())


def open_file(path, mode: str = "r"):
    return MultiStorageClientFeature.import_package().open(path, mode)
