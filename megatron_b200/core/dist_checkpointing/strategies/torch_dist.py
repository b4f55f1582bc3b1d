"""``torch_dist`` save/load strategy on top of ``torch.distributed.checkpoint`` (DCP).

On-disk layout matches the reference's default format (``strategies/torch.py:588,852``):
``<dir>/.metadata`` (DCP global metadata), ``__{rank}_{n}.distcp`` data files, ``common.pt`` and
``metadata.json``.  Instead of converting to PyTorch ``ShardedTensor`` objects first (reference
``mcore_to_pyt_state_dict``) the planners below emit DCP ``WriteItem``/``ReadItem``s directly
from the ``(global_shape, global_offset, local_shape)`` triplet of each mcore ShardedTensor.
"""
from __future__ import annotations

import io
from typing import Optional, Any, Dict, List, Tuple

import torch
import torch.distributed as dist
from torch.distributed.checkpoint import FileSystemReader, FileSystemWriter
from torch.distributed.checkpoint.default_planner import DefaultLoadPlanner, DefaultSavePlanner
from torch.distributed.checkpoint.metadata import ChunkStorageMetadata, Metadata, MetadataIndex, TensorProperties, TensorStorageMetadata
from torch.distributed.checkpoint.planner import LoadPlan, ReadItem, SavePlan, TensorWriteData, WriteItem, WriteItemType
from torch.distributed.checkpoint.planner_helpers import create_read_items_for_chunk_list
from torch.distributed.checkpoint.state_dict_loader import load as dcp_load
from torch.distributed.checkpoint.state_dict_saver import save as dcp_save

from ..mapping import ShardedObject, ShardedTensor, is_main_replica


def _full_shape(st: ShardedTensor) -> Tuple[int, ...]:
    return (1,) * st.prepend_axis_num + tuple(st.local_shape)


def _view(st: ShardedTensor) -> torch.Tensor:
    return st.data.detach().reshape(_full_shape(st))


class MCoreSavePlanner(DefaultSavePlanner):
    """Every main-replica ShardedTensor becomes one SHARD write item; ShardedObjects become BYTE_IO."""

    def __init__(self, sharded_tensors: List[ShardedTensor], sharded_objects: List[ShardedObject], **kw):
        super().__init__()
        self._sts = [s for s in sharded_tensors if is_main_replica(s.replica_id)]
        self._sos = [s for s in sharded_objects if is_main_replica(s.replica_id)]
        self._by_index: Dict[MetadataIndex, Any] = {}

    def set_up_planner(self, state_dict=None, storage_meta=None, is_coordinator=False, **kw):
        self.is_coordinator = is_coordinator
        self.state_dict = {}

    def create_local_plan(self) -> SavePlan:
        items = []
        for st in self._sts:
            idx = MetadataIndex(st.key, torch.Size(st.global_offset))
            self._by_index[idx] = st
            items.append(WriteItem(
                index=idx, type=WriteItemType.SHARD,
                tensor_data=TensorWriteData(
                    chunk=ChunkStorageMetadata(offsets=torch.Size(st.global_offset), sizes=torch.Size(_full_shape(st))),
                    properties=TensorProperties(dtype=st.dtype), size=torch.Size(st.global_shape)),
            ))
        for so in self._sos:
            idx = MetadataIndex(so.unique_key)
            self._by_index[idx] = so
            items.append(WriteItem(index=idx, type=WriteItemType.BYTE_IO))
        self.plan = SavePlan(items, planner_data={})
        return self.plan

    def resolve_data(self, write_item: WriteItem):
        obj = self._by_index[write_item.index]
        if isinstance(obj, ShardedObject):
            # on-disk format of a ShardedObject (reference strategies/torch.py:317-322): a pickled LIST of the data of every local object with this key
            buf = io.BytesIO()
            torch.save([obj.data], buf)
            buf.seek(0)
            return buf
        return _view(obj).contiguous()


class MCoreLoadPlanner(DefaultLoadPlanner):
    def __init__(self, sharded_tensors: List[ShardedTensor], sharded_objects: List[ShardedObject]):
        super().__init__()
        self._sts, self._sos = sharded_tensors, sharded_objects
        self._dst: Dict[Tuple[str, Tuple[int, ...]], ShardedTensor] = {}
        self._objs: Dict[str, ShardedObject] = {}
        self.loaded_objects: Dict[str, Any] = {}

    def set_up_planner(self, state_dict, metadata: Metadata = None, is_coordinator: bool = False):
        self.metadata, self.is_coordinator, self.state_dict = metadata, is_coordinator, {}

    def create_local_plan(self) -> LoadPlan:
        reqs: List[ReadItem] = []
        md = self.metadata.state_dict_metadata
        for st in self._sts:
            if st.key not in md:
                raise KeyError(f"{st.key} not found in the checkpoint")
            smd = md[st.key]
            stored = tuple(smd.size)
            want = tuple(st.global_shape)
            if stored != want and not st.allow_shape_mismatch:
                raise ValueError(f"global shape mismatch for {st.key}: checkpoint {stored} vs requested {want}")
            chunk = ChunkStorageMetadata(offsets=torch.Size(st.global_offset), sizes=torch.Size(_full_shape(st)))
            items = create_read_items_for_chunk_list(st.key, smd, [chunk])
            for it in items:
                # several requested shards may share an fqn: make the destination resolvable
                self._dst[(st.key, tuple(it.dest_index.offset))] = st
            reqs += items
        for so in self._sos:
            if so.unique_key not in md:
                raise KeyError(f"{so.unique_key} not found in the checkpoint")
            self._objs[so.unique_key] = so
            reqs.append(ReadItem(type=LoadItemTypeBYTE, dest_index=MetadataIndex(so.unique_key), dest_offsets=torch.Size([0]),
                                 storage_index=MetadataIndex(so.unique_key), storage_offsets=torch.Size([0]), lengths=torch.Size([0])))
        return LoadPlan(reqs)

    def create_global_plan(self, global_plan):
        return global_plan

    def finish_plan(self, new_plan):
        return new_plan

    def load_bytes(self, read_item: ReadItem, value: io.BytesIO) -> None:
        payload = torch.load(value, weights_only=False)
        # reference format: [data] (one entry per local object of that key)
        self.loaded_objects[read_item.dest_index.fqn] = payload[0] if isinstance(payload, list) and len(payload) == 1 else payload

    def resolve_tensor(self, read_item: ReadItem) -> torch.Tensor:
        st = self._dst[(read_item.dest_index.fqn, tuple(read_item.dest_index.offset))]
        t = _view(st)
        for d, (off, ln) in enumerate(zip(read_item.dest_offsets, read_item.lengths)):
            t = t.narrow(d, off, ln)
        return t

    def commit_tensor(self, read_item: ReadItem, tensor: torch.Tensor) -> None:
        pass


from torch.distributed.checkpoint.planner import LoadItemType as _LIT  # noqa: E402

LoadItemTypeBYTE = _LIT.BYTE_IO


def _no_dist(process_group) -> bool:
    return not (dist.is_available() and dist.is_initialized())


def save_sharded(sharded_tensors: List[ShardedTensor], sharded_objects: List[ShardedObject], checkpoint_dir: str, process_group=None, thread_count: int = 2):
    planner = MCoreSavePlanner(sharded_tensors, sharded_objects)
    writer = FileSystemWriter(checkpoint_dir, thread_count=thread_count, sync_files=False)
    dcp_save({}, storage_writer=writer, planner=planner, process_group=process_group, no_dist=_no_dist(process_group))


def load_sharded(sharded_tensors: List[ShardedTensor], sharded_objects: List[ShardedObject], checkpoint_dir: str, process_group=None, no_dist: bool = False) -> Dict[str, Any]:
    """Fills ``st.data`` of every requested ShardedTensor in place; returns {unique_key: object}."""
    for st in sharded_tensors:
        if st.data is None:
            st.init_data(device="cpu")
    planner = MCoreLoadPlanner(sharded_tensors, sharded_objects)
    dcp_load({}, storage_reader=FileSystemReader(checkpoint_dir), planner=planner, process_group=process_group, no_dist=no_dist or _no_dist(process_group))
    return planner.loaded_objects


def load_tensors_metadata(checkpoint_dir: str) -> Dict[str, ShardedTensor]:
    """Global shapes/dtypes of everything in a checkpoint, without data."""
    md = FileSystemReader(checkpoint_dir).read_metadata()
    out = {}
    for k, v in md.state_dict_metadata.items():
        if isinstance(v, TensorStorageMetadata):
            shape = tuple(v.size)
            out[k] = ShardedTensor(k, None, v.properties.dtype, shape, shape, (0,) * len(shape), (1,) * len(shape))
    return out

# ---- save split into plan (collective, caller) / write (any process) / commit (collective, caller) -------------------------------------------------------
# Reference: strategies/filesystem_async.py:114-440 + async_utils.py:237-349 — the planning and the metadata commit are collective and stay on the training
# process; the file writing (the slow part) runs in a persistent worker PROCESS fed with host copies of the shards in shared memory.


class _ResolvedPlanner(DefaultSavePlanner):
    """Planner stand-in for the writer process: every write item's payload has been resolved (host tensors in shared memory / raw bytes) before the hand-off."""

    def __init__(self, payloads: Dict[MetadataIndex, Any]):
        super().__init__()
        self._payloads = payloads

    def resolve_data(self, write_item: WriteItem):
        d = self._payloads[write_item.index]
        return io.BytesIO(d) if isinstance(d, (bytes, bytearray)) else d


class SavePlanCache:
    """Save plans survive from one checkpoint to the next (reference ``strategies/torch.py:700-790``: cached central plan / local plan / global metadata).

    The expensive part of planning is collective: every rank's local plan is pickled to the coordinator, merged, de-duplicated and scattered back.  A training
    run saves the SAME structure every time, so each rank fingerprints its local plan (keys, offsets, shapes, dtypes, object keys); if every rank's fingerprint
    matches its cached one — agreed with ONE integer all-reduce — the cached final plan and (on the coordinator) the cached global metadata are reused."""

    def __init__(self):
        self.fingerprint = None
        self.final_plan = None
        self.metadata = None
        self.hits = 0
        self.misses = 0

    @staticmethod
    def fingerprint_of(local_plan) -> int:
        import hashlib

        h = hashlib.sha1()
        for it in local_plan.items:
            idx = it.index
            h.update(repr((idx.fqn, tuple(idx.offset) if idx.offset is not None else None, int(it.type.value) if hasattr(it.type, "value") else str(it.type))).encode())
            td = getattr(it, "tensor_data", None)
            if td is not None:
                h.update(repr((tuple(td.size), str(td.properties.dtype), tuple(td.chunk.offsets), tuple(td.chunk.sizes))).encode())
        return int.from_bytes(h.digest()[:7], "little")

    def lookup(self, fp: int, process_group, no_dist: bool) -> bool:
        mine = self.fingerprint is not None and self.fingerprint == fp
        if no_dist:
            ok = mine
        else:
            dev = "cuda" if dist.get_backend(process_group) == "nccl" else "cpu"
            t = torch.tensor([1 if mine else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=process_group)
            ok = bool(t.item())
        if ok:
            self.hits += 1
        else:
            self.misses += 1
        return ok

    def store(self, fp: int, final_plan, metadata) -> None:
        self.fingerprint, self.final_plan, self.metadata = fp, final_plan, metadata


_PLAN_CACHES: Dict[Any, SavePlanCache] = {}


def get_plan_cache(process_group=None) -> SavePlanCache:
    """One cache per process group (``None`` = the world)."""
    return _PLAN_CACHES.setdefault(id(process_group) if process_group is not None else None, SavePlanCache())


def plan_save(sharded_tensors: List[ShardedTensor], sharded_objects: List[ShardedObject], checkpoint_dir: str, process_group=None, cache: Optional[SavePlanCache] = None):
    """Collective planning.  Returns ``(final_plan, payloads, global_metadata_or_None)``; payload tensors are host copies (shared memory), detached from the
    training state, so training may overwrite the originals as soon as this returns.  With ``cache`` an unchanged structure skips the plan gather / merge /
    scatter (see ``SavePlanCache``)."""
    import copy

    no_dist = _no_dist(process_group)
    planner = MCoreSavePlanner(sharded_tensors, sharded_objects)
    writer = FileSystemWriter(checkpoint_dir, sync_files=False)
    rank = 0 if no_dist else dist.get_rank(process_group)
    world = 1 if no_dist else dist.get_world_size(process_group)
    coordinator = rank == 0
    planner.set_up_planner({}, None, coordinator)
    writer.set_up_storage_writer(coordinator)
    local_plan = writer.prepare_local_plan(planner.create_local_plan())
    fp = SavePlanCache.fingerprint_of(local_plan) if cache is not None else None
    if cache is not None and cache.lookup(fp, process_group, no_dist):
        final_plan, metadata = cache.final_plan, (copy.deepcopy(cache.metadata) if cache.metadata is not None else None)
    else:
        if no_dist:
            plans = [local_plan]
        else:
            plans = [None] * world
            dist.all_gather_object(plans, local_plan, group=process_group)
        metadata = None
        if coordinator:
            plans, metadata = planner.create_global_plan(plans)
            plans = writer.prepare_global_plan(plans)
        if no_dist:
            final_plan = plans[0]
        else:
            out = [None]
            dist.scatter_object_list(out, plans if coordinator else None, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            final_plan = out[0]
        final_plan = planner.finish_plan(final_plan)
        if cache is not None:
            cache.store(fp, final_plan, copy.deepcopy(metadata) if metadata is not None else None)
    payloads: Dict[MetadataIndex, Any] = {}
    for item in final_plan.items:
        data = planner.resolve_data(item)
        if isinstance(data, io.BytesIO):
            payloads[item.index] = data.getvalue()
        else:
            host = data.detach().to("cpu", copy=True).contiguous()
            payloads[item.index] = host.share_memory_()
    return final_plan, payloads, metadata


def write_planned(final_plan, payloads, checkpoint_dir: str, thread_count: int = 2):
    """The file writing of one rank; safe to run in another process (no process group, no CUDA).  Returns the WriteResults."""
    writer = FileSystemWriter(checkpoint_dir, thread_count=thread_count, sync_files=False)
    writer.set_up_storage_writer(False)
    fut = writer.write_data(final_plan, _ResolvedPlanner(payloads))
    fut.wait()
    return fut.value()


def commit_save(results, metadata, checkpoint_dir: str, process_group=None):
    """Collective: gather every rank's WriteResults; the coordinator writes ``.metadata``."""
    no_dist = _no_dist(process_group)
    if no_dist:
        all_results = [results]
    else:
        world = dist.get_world_size(process_group)
        all_results = [None] * world
        dist.all_gather_object(all_results, results, group=process_group)
    if metadata is not None:
        writer = FileSystemWriter(checkpoint_dir, sync_files=False)
        writer.set_up_storage_writer(True)
        writer.finish(metadata, all_results)
