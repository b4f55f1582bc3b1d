"""Sharded state dict ⇄ (flat tensor list, hollow skeleton): the interface a LOCAL checkpoint manager needs (reference
``dist_checkpointing/tensor_aware_state_dict.py:48-400``, built for nvidia_resiliency_ext's ``LocalCheckpointManager``).

A local checkpoint never reshards, so nothing has to be planned or described globally: the manager only wants the bulk tensors (to copy to host /
node-local storage / a peer asynchronously) and a small picklable remainder.  ``from_state_dict`` splits a sharded state dict into

* ``common``  – everything that is not sharded (iteration, args, scheduler state …),
* ``sharded`` – the ShardedTensor / ShardedObject leaves; with ``algo='fully_parallel'`` each DP-replicated tensor is kept ONLY by the rank that is its main
  replica inside ``parallelization_group`` (the others re-receive it by broadcast when the state dict is rebuilt), ``'atomic'`` keeps everything everywhere.

``pop_tensors`` hollows the container (tensor data → shape / dtype / device records), ``insert_tensors`` refills it, ``to_state_dict`` writes the stored data into
a freshly generated sharded state dict of the SAME layout and returns the plain state dict ``dist_checkpointing.load`` would have returned."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Iterable, Iterator, List, Optional

import torch
import torch.distributed as dist

from .dict_utils import dict_list_map_inplace, extract_matching_values, merge, nested_values
from .mapping import LocalNonpersistentObject, ShardedBase, ShardedObject, ShardedTensor, ShardedTensorFactory, apply_factories, apply_factory_merges, is_main_replica


@dataclass
class _Hollow:
    shape: torch.Size
    dtype: torch.dtype
    device: torch.device


def _replica_in_group(st: ShardedTensor, group) -> bool:
    """Is this rank the writer of ``st`` among the ranks of ``group``?  The last component of the replica id is the data-parallel index."""
    rid = st.replica_id
    return (rid[-1] if isinstance(rid, tuple) else rid) == 0


class MCoreTensorAwareStateDict:
    def __init__(self, common: Dict, sharded: Dict, algo: str, group):
        self.common, self.sharded, self.algo, self.group = common, sharded, algo, group
        self._dropped: List[str] = []           # keys of replicated tensors this rank does not store (fully_parallel)

    # ---- construction ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _validate(algo: str) -> None:
        if algo not in ("atomic", "fully_parallel"):
            raise NotImplementedError(f"only the 'atomic' and 'fully_parallel' algorithms are supported, got {algo!r}")

    @classmethod
    def from_state_dict(cls, sharded_state_dict: Dict, algo: str = "atomic", parallelization_group=None) -> "MCoreTensorAwareStateDict":
        cls._validate(algo)
        sharded, rest = extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedBase))
        _, common = extract_matching_values(rest, lambda v: isinstance(v, LocalNonpersistentObject))
        apply_factories(sharded)
        self = cls(common, sharded, algo, parallelization_group)
        if algo == "fully_parallel":
            for st in self._sharded_tensors():
                if not _replica_in_group(st, parallelization_group):
                    self._dropped.append(st.key)
                    st.data = _Hollow(st.data.shape, st.data.dtype, st.data.device)
        return self

    # ---- views -------------------------------------------------------------------------------------------------------------------
    def _sharded_tensors(self) -> List[ShardedTensor]:
        return [v for v in nested_values(self.sharded) if isinstance(v, ShardedTensor)]

    def _stored(self) -> List[ShardedTensor]:
        return [st for st in self._sharded_tensors() if st.key not in self._dropped or isinstance(st.data, torch.Tensor)]

    @property
    def is_hollow(self) -> bool:
        stored = self._stored()
        return bool(stored) and all(isinstance(st.data, _Hollow) for st in stored if st.key not in self._dropped) and any(st.key not in self._dropped for st in stored)

    @property
    def tensors(self) -> Iterator[torch.Tensor]:
        assert not self.is_hollow, "the tensors were popped"
        return (st.data for st in self._sharded_tensors() if isinstance(st.data, torch.Tensor))

    @property
    def common_state_dict(self) -> Dict:
        return self.common

    # ---- hollow / refill ---------------------------------------------------------------------------------------------------------
    def pop_tensors(self) -> List[torch.Tensor]:
        out = []
        for st in self._sharded_tensors():
            if isinstance(st.data, torch.Tensor):
                out.append(st.data)
                st.data = _Hollow(st.data.shape, st.data.dtype, st.data.device)
        return out

    def insert_tensors(self, tensor_data: Iterable[torch.Tensor]) -> None:
        slots = [st for st in self._sharded_tensors() if st.key not in self._dropped]
        data = list(tensor_data)
        if len(data) != len(slots):
            raise ValueError(f"expected {len(slots)} tensors, got {len(data)}")
        for st, t in zip(slots, data):
            rec = st.data
            if isinstance(rec, _Hollow) and (tuple(t.shape) != tuple(rec.shape) or t.dtype != rec.dtype):
                raise ValueError(f"tensor for {st.key}: expected {tuple(rec.shape)} {rec.dtype}, got {tuple(t.shape)} {t.dtype}")
            st.data = t

    def init_tensors(self) -> None:
        """Allocate (uninitialised) storage for every hollow slot — the receive buffers of a peer-to-peer refill."""
        for st in self._sharded_tensors():
            if isinstance(st.data, _Hollow) and st.key not in self._dropped:
                st.data = torch.empty(st.data.shape, dtype=st.data.dtype, device=st.data.device)

    def copy_tensors_to_cpu(self, non_blocking: bool = False) -> None:
        for st in self._sharded_tensors():
            if isinstance(st.data, torch.Tensor):
                dev = st.data.device
                host = torch.empty(st.data.shape, dtype=st.data.dtype, device="cpu", pin_memory=dev.type == "cuda")
                host.copy_(st.data, non_blocking=non_blocking)
                host._orig_device = dev
                st.data = host

    def restore_tensor_device(self, non_blocking: bool = True) -> None:
        for st in self._sharded_tensors():
            if isinstance(st.data, torch.Tensor) and getattr(st.data, "_orig_device", None) is not None:
                st.data = st.data.to(st.data._orig_device, non_blocking=non_blocking)

    # ---- back to a state dict --------------------------------------------------------------------------------------------------------
    def to_state_dict(self, sharded_state_dict: Dict, algo: Optional[str] = None, parallelization_group=None) -> Dict:
        """Fill a freshly generated sharded state dict (same parallel layout as at save time) from the stored data → plain state dict."""
        algo = algo or self.algo
        self._validate(algo)
        group = parallelization_group if parallelization_group is not None else self.group
        assert not self.is_hollow, "insert_tensors() first"
        have = {st.key + repr(st.global_offset): st for st in self._sharded_tensors()}
        have_obj = {v.unique_key: v for v in nested_values(self.sharded) if isinstance(v, ShardedObject)}
        template, _ = extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedTensorFactory), return_lists_as_dicts=True)
        sharded, rest = extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedBase))
        nonpers, _ = extract_matching_values(rest, lambda v: isinstance(v, LocalNonpersistentObject))
        apply_factories(sharded)
        wanted = [v for v in nested_values(sharded) if isinstance(v, ShardedTensor)]
        for st in wanted:
            src = have.get(st.key + repr(st.global_offset))
            if src is None:
                raise KeyError(f"local checkpoint holds no shard {st.key} at offset {st.global_offset} (the parallel layout changed: use the global checkpoint)")
            if isinstance(src.data, torch.Tensor):
                if st.data is not None and isinstance(st.data, torch.Tensor):
                    st.data.copy_(src.data)
                else:
                    st.data = src.data
        if algo == "fully_parallel" and dist.is_initialized() and dist.get_world_size(group) > 1:
            # tensors replicated across the group live on its first rank only: agree on the list (union of what the others dropped), then broadcast in key order
            lists = [None] * dist.get_world_size(group)
            dist.all_gather_object(lists, sorted(self._dropped), group=group)
            replicated = sorted(set().union(*lists))
            src_rank = dist.get_global_rank(group, 0) if group is not None else 0
            by_key: Dict[str, List[ShardedTensor]] = {}
            for st in wanted:
                by_key.setdefault(st.key, []).append(st)
            for key in replicated:
                for st in sorted(by_key.get(key, []), key=lambda x: x.global_offset):
                    if not isinstance(st.data, torch.Tensor):
                        rec = have[st.key + repr(st.global_offset)].data
                        st.data = torch.empty(rec.shape, dtype=rec.dtype, device=rec.device)
                    dist.broadcast(st.data, src=src_rank, group=group)

        def unwrap(v):
            if isinstance(v, ShardedTensor):
                return v.data
            if isinstance(v, ShardedObject):
                return have_obj[v.unique_key].data if v.unique_key in have_obj else v.data
            return v

        dict_list_map_inplace(unwrap, sharded)
        if template:
            t, _ = extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedTensorFactory))
            sharded = apply_factory_merges(sharded, t)
        out = dict(self.common)
        merge(out, sharded)
        dict_list_map_inplace(lambda o: o.unwrap() if isinstance(o, LocalNonpersistentObject) else o, nonpers)
        if nonpers:
            merge(out, nonpers)
        return out
