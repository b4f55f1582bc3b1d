"""Registry of NVLink collective back ends keyed by process group.

``backend_for(group)`` returns the :class:`megatron_b200.parallel.nvlink.NVLinkBackend`
bound to ``group`` when the symmetric-memory runtime has been enabled for it
(``enable_for_group``), else ``None`` so callers fall back to ``torch.distributed``.
"""
from __future__ import annotations

from typing import Dict, Optional

_BACKENDS: Dict[int, object] = {}
_DISABLED = False


def backend_for(group) -> Optional[object]:
    if _DISABLED or group is None:
        return None
    return _BACKENDS.get(id(group))


def enable_for_group(group, **kwargs):
    """Create (collectively) the symmetric heap + NVLink kernels for ``group``."""
    from .nvlink import NVLinkBackend

    be = _BACKENDS.get(id(group))
    if be is None:
        from .nvlink_debug import maybe_wrap

        be = maybe_wrap(NVLinkBackend(group, **kwargs), group)      # MEGATRON_B200_NVL_DEBUG=1: every collective is cross-checked against NCCL
        _BACKENDS[id(group)] = be
    return be


def disable_all(flag: bool = True):
    global _DISABLED
    _DISABLED = flag


def reset():
    _BACKENDS.clear()
