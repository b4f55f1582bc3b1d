"""Key/prefix manipulation over sharded state dicts (reference ``dist_checkpointing/utils.py``)."""
from __future__ import annotations

from typing import Dict, Tuple

from .dict_utils import dict_list_map_inplace, extract_matching_values
from .mapping import LocalNonpersistentObject, ShardedBase, ShardedObject, ShardedStateDict, ShardedTensor, ShardedTensorFactory, StateDict


def extract_sharded_tensors(sd) -> Tuple[ShardedStateDict, StateDict]:
    return extract_matching_values(sd, lambda v: isinstance(v, ShardedTensor))


def extract_sharded_tensors_and_factories(sd):
    return extract_matching_values(sd, lambda v: isinstance(v, (ShardedTensor, ShardedTensorFactory)))


def extract_sharded_base(sd):
    return extract_matching_values(sd, lambda v: isinstance(v, ShardedBase))


def extract_nonpersistent(sd):
    return extract_matching_values(sd, lambda v: isinstance(v, LocalNonpersistentObject))


def add_prefix_for_sharding(sharded_state_dict: ShardedStateDict, prefix: str):
    def f(t):
        if isinstance(t, (ShardedTensor, ShardedTensorFactory, ShardedObject)):
            t.key = f"{prefix}{t.key}"
        return t

    dict_list_map_inplace(f, sharded_state_dict)


def replace_prefix_for_sharding(sharded_state_dict: ShardedStateDict, old_prefix: str, new_prefix: str):
    def f(x):
        if isinstance(x, (ShardedTensor, ShardedTensorFactory, ShardedObject)):
            if not x.key.startswith(old_prefix):
                raise ValueError(f"expected {x.key} to begin with prefix {old_prefix}")
            x.key = f"{new_prefix}{x.key[len(old_prefix):]}"
        return x

    dict_list_map_inplace(f, sharded_state_dict)


def apply_prefix_mapping(sharded_state_dict: ShardedStateDict, prefix_map: Dict[str, str]):
    def f(x):
        if isinstance(x, (ShardedTensor, ShardedTensorFactory, ShardedObject)):
            for old, new in prefix_map.items():
                if x.key.startswith(old):
                    x.key = f"{new}{x.key[len(old):]}"
                    break
        return x

    dict_list_map_inplace(f, sharded_state_dict)
