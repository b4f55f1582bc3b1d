"""Map optimizer tensors onto the sharding of their model parameter
(reference ``dist_checkpointing/optimizer.py:83-150``)."""
from __future__ import annotations

from dataclasses import replace
from typing import Dict, Iterable, Union

import torch

from .dict_utils import nested_values
from .mapping import ShardedStateDict, ShardedTensor, ShardedTensorFactory


def get_param_id_to_sharded_param_map(model_sharded_state_dict: ShardedStateDict, optim_params_iter: Iterable[torch.nn.Parameter]) -> Dict[int, Union[ShardedTensor, ShardedTensorFactory]]:
    """Match optimizer params to model ShardedTensors by tensor identity (data_ptr + shape)."""
    by_id = {}
    for sh in nested_values(model_sharded_state_dict):
        if isinstance(sh, (ShardedTensor, ShardedTensorFactory)) and sh.data is not None:
            by_id[id(sh.data)] = sh
            by_id[(sh.data.data_ptr(), tuple(sh.data.shape))] = sh
    out = {}
    for i, p in enumerate(optim_params_iter):
        sh = by_id.get(id(p)) or by_id.get((p.data_ptr(), tuple(p.shape)))
        if sh is None:
            raise KeyError(f"optimizer parameter #{i} of shape {tuple(p.shape)} not found in the model sharded state dict")
        out[i] = sh
    return out


def make_sharded_optimizer_tensor(model_param: Union[ShardedTensor, ShardedTensorFactory], optim_param: torch.Tensor, prefix: str, replica_id=None):
    """Same global shape/offsets as the model parameter, different key prefix and data/dtype."""
    if isinstance(model_param, ShardedTensorFactory):
        return replace(model_param, key=f"{prefix}.{model_param.key}", data=optim_param)
    assert tuple(optim_param.shape) == tuple(model_param.local_shape), f"optimizer tensor shape {tuple(optim_param.shape)} != model shard {model_param.local_shape} for {model_param.key}"
    return replace(model_param, key=f"{prefix}.{model_param.key}", data=optim_param, dtype=optim_param.dtype,
                   replica_id=model_param.replica_id if replica_id is None else replica_id)
