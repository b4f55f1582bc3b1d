"""Transformer helpers: sharded-state-dict builders, attention mask utils
(reference ``transformer/utils.py:96,252``)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import parallel_state as ps
from ..dist_checkpointing.mapping import ShardedObject, ShardedStateDict, StateDict
from ..utils import make_sharded_tensor_for_checkpoint, make_tp_sharded_tensor_for_checkpoint


def get_linear_layer(rows, columns, init_method, perform_initialization=True):
    layer = torch.nn.Linear(rows, columns)
    if perform_initialization:
        init_method(layer.weight)
    with torch.no_grad():
        layer.bias.zero_()
    return layer


def get_default_causal_mask(sq: int, device=None) -> torch.Tensor:
    return torch.triu(torch.ones(sq, sq, device=device), diagonal=1).bool()


def attention_mask_func(attention_scores, attention_mask):
    return attention_scores.masked_fill(attention_mask, -10000.0)


def make_sharded_tensors_for_checkpoint(
    state_dict: StateDict,
    prefix: str,
    tensor_parallel_layers_axis_map: Optional[Dict[str, int]] = None,
    sharded_offsets: Iterable[Tuple[int, int, int]] = (),
    extra_state_suffix: str = "_extra_state",
    tp_group=None,
    dp_cp_group=None,
) -> ShardedStateDict:
    """Wrap every tensor of ``state_dict``: keys in the axis map are TP-sharded along that
    axis, all others are TP-replicated; ``*_extra_state`` entries become ShardedObjects."""
    axis_map = tensor_parallel_layers_axis_map or {}
    out = {}
    for name, t in state_dict.items():
        key = f"{prefix}{name}"
        if name.endswith(extra_state_suffix):
            out[key] = make_sharded_object_for_checkpoint(t, key, sharded_offsets)
        elif name in axis_map:
            out[key] = make_tp_sharded_tensor_for_checkpoint(t, key, axis_map[name], prepend_offsets=sharded_offsets, tp_group=tp_group, dp_cp_group=dp_cp_group)
        else:
            out[key] = make_sharded_tensor_for_checkpoint(t, key, prepend_offsets=sharded_offsets, tp_group=tp_group, dp_cp_group=dp_cp_group)
    return out


def make_sharded_object_for_checkpoint(obj, key: str, sharded_offsets=(), replica_id=None, **kwargs):
    if replica_id is None:
        replica_id = (0, ps.get_tensor_model_parallel_rank(), ps.get_data_parallel_rank(with_context_parallel=True))
    return ShardedObject(key, obj, *_get_extra_state_offsets(sharded_offsets), replica_id, **kwargs)


def _get_extra_state_offsets(sharded_offsets):
    if sharded_offsets:
        so = sorted(sharded_offsets, key=lambda x: x[0])
        axis, off, shape = zip(*so)
        assert list(axis) == list(range(len(axis))), f"expected contiguous axis for offsets: {so}"
        return tuple(shape), tuple(off)
    return (1,), (0,)


def sharded_state_dict_default(module, prefix="", sharded_offsets=(), metadata=None, tp_group=None):
    """Use ``module.sharded_state_dict`` when it exists, else treat all tensors as replicated."""
    if hasattr(module, "sharded_state_dict"):
        return module.sharded_state_dict(prefix=prefix, sharded_offsets=sharded_offsets, metadata=metadata)
    sd = module.state_dict(prefix="", keep_vars=True)
    return make_sharded_tensors_for_checkpoint(sd, prefix, {}, sharded_offsets, tp_group=tp_group)


def ensure_metadata_has_dp_cp_group(metadata):
    metadata = dict(metadata or {})
    if "dp_cp_group" not in metadata:
        try:
            metadata["dp_cp_group"] = ps.get_data_parallel_group(with_context_parallel=True)
        except AssertionError:
            metadata["dp_cp_group"] = None
    return metadata
