"""Allow-lists for checkpoint deserialisation (reference ``safe_globals.py:54-125``).

Checkpoints are loaded with ``torch.load(weights_only=True)``; the classes our own checkpoints legitimately contain are registered
with torch's safe-globals list, and the two places that still meet raw pickles (per-module ``_extra_state`` blobs, ``numpy.load`` of
legacy index files) go through an unpickler that can only build plain containers, scalars and tensors — a crafted file cannot name
an arbitrary callable."""
from __future__ import annotations

import io
import pickle
import threading
from argparse import Namespace
from collections import OrderedDict
from unittest.mock import patch

import numpy
import torch

from .enums import ModelType
from .rerun_state_machine import RerunDiagnostic, RerunMode, RerunState
from .enums import AttnBackend

SAFE_GLOBALS = [Namespace, OrderedDict, ModelType, AttnBackend, RerunDiagnostic, RerunMode, RerunState, numpy.dtype, numpy.ndarray,
                type(numpy.dtype("int64")), type(numpy.dtype("float32")), type(numpy.dtype("uint8"))]
try:                                        # numpy ≥ 2 moved the reconstruct helpers
    from numpy._core.multiarray import _reconstruct as _np_reconstruct, scalar as _np_scalar
except ImportError:                         # pragma: no cover
    from numpy.core.multiarray import _reconstruct as _np_reconstruct, scalar as _np_scalar
SAFE_GLOBALS += [_np_reconstruct, _np_scalar]

_pickle_patch_lock = threading.Lock()


def register_safe_globals() -> None:
    torch.serialization.add_safe_globals(list(SAFE_GLOBALS))


def safe_load_from_bytes(b: bytes):
    return torch.load(io.BytesIO(b), weights_only=True)


class SafeUnpickler(pickle.Unpickler):
    _SAFE_CLASSES = frozenset({
        ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "int"), ("builtins", "float"),
        ("builtins", "bool"), ("builtins", "bytes"), ("builtins", "str"), ("builtins", "complex"), ("collections", "OrderedDict"),
        ("torch", "Size"), ("torch", "device"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._tensor", "_rebuild_from_type_v2"),
        ("torch.storage", "UntypedStorage"), ("numpy", "dtype"), ("numpy", "ndarray"),
        ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "_reconstruct"),
        ("numpy._core.multiarray", "scalar"), ("numpy.core.multiarray", "scalar"),
        ("megatron_b200.core.safe_globals", "safe_load_from_bytes"),
    })
    _SAFE_MODULE_ATTR_PREFIXES = (("torch", ("float", "bfloat", "int", "uint", "bool", "complex", "half", "double", "long", "short")),)

    def find_class(self, module: str, name: str):
        # (the reference package, when imported in the same process, re-points torch's tensor pickling at ITS reader: same meaning)
        if (module, name) in (("torch.storage", "_load_from_bytes"), ("megatron.core.safe_globals", "safe_load_from_bytes")):
            return safe_load_from_bytes                                  # the weights_only reader instead of a nested full unpickle
        if (module, name) in self._SAFE_CLASSES:
            return super().find_class(module, name)
        for mod, prefixes in self._SAFE_MODULE_ATTR_PREFIXES:
            if module == mod and name.startswith(prefixes) and isinstance(getattr(torch, name, None), torch.dtype):
                return getattr(torch, name)
        raise pickle.UnpicklingError(f"checkpoint references {module}.{name}, which is not on the allow-list")


def _safe_pickle_load(file, **kwargs):
    return SafeUnpickler(file, **{k: v for k, v in kwargs.items() if k in ("fix_imports", "encoding", "errors", "buffers")}).load()


def safe_pickle_loads(data: bytes):
    return SafeUnpickler(io.BytesIO(data)).load()


def safe_numpy_load(path, **kwargs):
    """``numpy.load(allow_pickle=True)`` with the restricted unpickler swapped in for object arrays."""
    with _pickle_patch_lock:
        with patch("pickle.load", _safe_pickle_load):
            return numpy.load(path, **kwargs)
