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


# ---- variable-count collectives (reference: Triton ``inference/communication/torch_symm_triton/variable_collectives.py``) ----------------------------------
def all_gather_v(x, sizes, group=None):
    """Concatenate per-rank tensors with DIFFERENT first-dim sizes (``sizes[r]`` rows on rank r; MoE token counts, ragged decode batches).
    NCCL handles uneven lists natively; backends that need equal shapes (gloo, the NVLink multimem kernels) go through one padded gather + a slice —
    the padding is at most ``max(sizes) - min(sizes)`` rows per rank."""
    import torch
    import torch.distributed as dist

    ws = dist.get_world_size(group)
    assert len(sizes) == ws and x.shape[0] == sizes[dist.get_rank(group)]
    tail = tuple(x.shape[1:])
    if dist.get_backend(group) == "nccl":
        outs = [x.new_empty((n,) + tail) for n in sizes]
        dist.all_gather(outs, x.contiguous(), group=group)
        return torch.cat(outs, dim=0)
    m = max(sizes)
    pad = x.new_zeros((m,) + tail)
    pad[: x.shape[0]] = x
    be = _BACKENDS.get(id(group)) if not _DISABLED else None
    if be is not None and x.is_cuda:
        full = be.all_gather(pad)
    else:
        full = x.new_empty((m * ws,) + tail)
        dist.all_gather_into_tensor(full, pad, group=group)
    return torch.cat([full[r * m : r * m + n] for r, n in enumerate(sizes)], dim=0)


def reduce_scatter_v(x, sizes, group=None):
    """Inverse: ``x`` holds ``sum(sizes)`` rows (rank-major); rank r receives the SUM over ranks of its ``sizes[r]`` rows."""
    import torch
    import torch.distributed as dist

    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    assert x.shape[0] == sum(sizes)
    tail = tuple(x.shape[1:])
    parts = list(torch.split(x, list(sizes), dim=0))
    if dist.get_backend(group) == "nccl":
        out = x.new_empty((sizes[rank],) + tail)
        dist.reduce_scatter(out, [p.contiguous() for p in parts], group=group)
        return out
    m = max(sizes)
    padded = x.new_zeros((ws * m,) + tail)
    for r, p in enumerate(parts):
        padded[r * m : r * m + p.shape[0]] = p
    out = x.new_empty((m,) + tail)
    dist.reduce_scatter_tensor(out, padded, group=group)
    return out[: sizes[rank]]
