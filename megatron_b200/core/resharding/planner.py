"""Plan weight movement between two parallel layouts (reference ``resharding/planner.py``; refit for RL / online TP-PP-EP change).

Every rank describes what it HOLDS (source shards) and what it NEEDS (destination shards) as axis-aligned boxes of named global
tensors — exactly the information in a ``ShardedTensor`` (key, global_shape, global_offset, local_shape).  After one all-gather of
these descriptors each rank derives, deterministically and identically, the list of box intersections to move.  Among replicated
sources the one on the destination rank itself is preferred (no traffic), then the replica with the fewest bytes scheduled so far
(balances NVLink egress)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass(frozen=True)
class ShardDesc:
    key: str
    global_shape: Tuple[int, ...]
    offset: Tuple[int, ...]
    shape: Tuple[int, ...]
    rank: int

    @staticmethod
    def from_sharded_tensor(sh, rank: int) -> "ShardDesc":
        return ShardDesc(sh.key, tuple(sh.global_shape), tuple(sh.global_offset), tuple(sh.local_shape), rank)


@dataclass(frozen=True)
class TransferOp:
    key: str
    src_rank: int
    dst_rank: int
    src_slices: Tuple[Tuple[int, int], ...]   # (start, stop) inside the source local tensor
    dst_slices: Tuple[Tuple[int, int], ...]   # inside the destination local tensor
    numel: int


def _intersect(a: ShardDesc, b: ShardDesc) -> Optional[Tuple[Tuple[int, int], ...]]:
    box = []
    for ao, asz, bo, bsz in zip(a.offset, a.shape, b.offset, b.shape):
        lo, hi = max(ao, bo), min(ao + asz, bo + bsz)
        if lo >= hi:
            return None
        box.append((lo, hi))
    return tuple(box)


def build_reshard_plan(sources: Sequence[ShardDesc], dests: Sequence[ShardDesc]) -> List[TransferOp]:
    by_key: Dict[str, List[ShardDesc]] = {}
    for s in sources:
        by_key.setdefault(s.key, []).append(s)
    egress: Dict[int, int] = {}
    ops: List[TransferOp] = []
    for d in sorted(dests, key=lambda x: (x.key, x.rank, x.offset)):
        cands = by_key.get(d.key)
        if not cands:
            raise KeyError(f"no source holds tensor '{d.key}' needed by rank {d.rank}")
        if tuple(cands[0].global_shape) != tuple(d.global_shape):
            raise ValueError(f"global shape mismatch for '{d.key}': {cands[0].global_shape} vs {d.global_shape}")
        # cover the destination box with disjoint pieces; `todo` holds the still-uncovered sub-boxes
        todo = [tuple((o, o + s) for o, s in zip(d.offset, d.shape))]
        for src in sorted(cands, key=lambda s: (s.rank != d.rank, egress.get(s.rank, 0), s.rank, s.offset)):
            nxt = []
            for box in todo:
                bdesc = ShardDesc(d.key, d.global_shape, tuple(b[0] for b in box), tuple(b[1] - b[0] for b in box), d.rank)
                inter = _intersect(src, bdesc)
                if inter is None:
                    nxt.append(box)
                    continue
                numel = 1
                for lo, hi in inter:
                    numel *= hi - lo
                ops.append(TransferOp(d.key, src.rank, d.rank, tuple((lo - so, hi - so) for (lo, hi), so in zip(inter, src.offset)),
                                      tuple((lo - do, hi - do) for (lo, hi), do in zip(inter, d.offset)), numel))
                egress[src.rank] = egress.get(src.rank, 0) + (numel if src.rank != d.rank else 0)
                # split the remainder of `box` around `inter` into axis-aligned pieces
                rem = list(box)
                for ax, ((lo, hi), (blo, bhi)) in enumerate(zip(inter, box)):
                    if blo < lo:
                        nxt.append(tuple(rem[:ax] + [(blo, lo)] + list(box[ax + 1:])))
                    if hi < bhi:
                        nxt.append(tuple(rem[:ax] + [(hi, bhi)] + list(box[ax + 1:])))
                    rem[ax] = (lo, hi)
            todo = nxt
            if not todo:
                break
        if todo:
            raise ValueError(f"sources do not cover '{d.key}' box(es) {todo} needed by rank {d.rank}")
    return ops


# =====================================================================================================================
# Model refit planner (reference ``resharding/planner.py:26-524``): ParameterMetadata rosters -> per-rank ReshardPlan
# =====================================================================================================================
from .utils import (  # noqa: E402
    ParameterMetadata, ReshardPlan, ShardingDescriptor, extract_module_metadata, select_src_metadata_balanced,
)
from .utils import TransferOp as RefitTransferOp  # noqa: E402


def _intersect_runs(a, b):
    """Pieces common to two run lists of one axis: [(a_local_start, b_local_start, length)], ordered by b's local offset."""
    out = []
    for al, ag, an in a:
        for bl, bg, bn in b:
            lo, hi = max(ag, bg), min(ag + an, bg + bn)
            if lo < hi:
                out.append((al + lo - ag, bl + lo - bg, hi - lo))
    out.sort(key=lambda t: t[1])
    # merge pieces that are adjacent on BOTH sides (a TP=2 -> TP=2 copy is one piece again, not `stride` of them)
    merged = []
    for p in out:
        if merged and merged[-1][0] + merged[-1][2] == p[0] and merged[-1][1] + merged[-1][2] == p[1]:
            merged[-1] = (merged[-1][0], merged[-1][1], merged[-1][2] + p[2])
        else:
            merged.append(p)
    return merged


def _build_descriptors_for_param(src: ParameterMetadata, dst: ParameterMetadata):
    """Which dimensions are sharded differently on the two sides (reference ``planner.py:31``; used for reports / validation)."""
    out = []
    if src.is_tp or dst.is_tp:
        dim = dst.partition_dim if dst.is_tp else src.partition_dim
        out.append(ShardingDescriptor("tp", dim, src.partition_stride if src.is_tp else 1, dst.partition_stride if dst.is_tp else 1,
                                      list(src.tensor_parallel_group_ranks or [src.owner_rank]) if src.is_tp else [src.owner_rank],
                                      list(dst.tensor_parallel_group_ranks or [dst.owner_rank]) if dst.is_tp else [dst.owner_rank]))
    if src.is_fused_experts() or dst.is_fused_experts():
        out.append(ShardingDescriptor("ep", 0, 1, 1, list(src.expert_parallel_group_ranks or [src.owner_rank]), list(dst.expert_parallel_group_ranks or [dst.owner_rank])))
    return out


def _shard_identity(m: ParameterMetadata):
    """Replicas (data-parallel copies) of one shard have the same identity."""
    return (m.resolved_name or m.name, m.runs())


def index_metadata_rosters(gathered_pairs):
    """``gathered_pairs[rank] = (src_metadata_list, dst_metadata_list)`` -> (sources by resolved name -> shard identity ->
    replicas, destinations by rank) (reference ``planner.py:388``)."""
    src_index, dst_by_rank = {}, {}
    for pair in gathered_pairs:
        if pair is None:
            continue
        for m in pair[0]:
            src_index.setdefault(m.resolved_name or m.name, {}).setdefault(_shard_identity(m), []).append(m)
        for m in pair[1]:
            dst_by_rank.setdefault(m.owner_rank, []).append(m)
    return src_index, dst_by_rank


def _iter_global_transfer_ops(src_index, dst_by_rank):
    """Deterministic global list of (task_id, name, src_meta, dst_meta, src_slices, dst_slices, numel): every rank derives the
    same list from the same rosters, so matching send / recv pairs share a ``task_id`` without any further communication."""
    task = 0
    for dst_rank in sorted(dst_by_rank):
        for d in sorted(dst_by_rank[dst_rank], key=lambda m: m.resolved_name or m.name):
            key = d.resolved_name or d.name
            shards = src_index.get(key)
            if not shards:
                raise KeyError(f"refit: no source rank holds '{key}' (needed by rank {dst_rank} as '{d.name}')")
            d_runs = d.runs()
            covered = 0
            for ident in sorted(shards, key=lambda i: i[1]):
                replicas = shards[ident]
                s0 = replicas[0]
                if s0.global_shape() != d.global_shape():
                    raise ValueError(f"refit: '{key}' has global shape {s0.global_shape()} at the source and {d.global_shape()} at the destination")
                per_axis = [_intersect_runs(sr, dr) for sr, dr in zip(ident[1], d_runs)]
                if any(not ax for ax in per_axis):
                    continue
                s = select_src_metadata_balanced(replicas, d, dst_rank)
                # cartesian product of the per-axis pieces (almost always 1 x ... x k x ... x 1)
                combos = [[]]
                for ax in per_axis:
                    combos = [c + [p] for c in combos for p in ax]
                for c in combos:
                    numel = 1
                    for p in c:
                        numel *= p[2]
                    covered += numel
                    yield (task, key, s, d, tuple(slice(p[0], p[0] + p[2]) for p in c), tuple(slice(p[1], p[1] + p[2]) for p in c), numel)
                    task += 1
            need = 1
            for n in d.shape:
                need *= n
            if covered != need:
                raise ValueError(f"refit: sources cover {covered} of {need} elements of '{key}' on rank {dst_rank}")


def build_plan_from_rosters(gathered_pairs, my_global_rank: int) -> ReshardPlan:
    """Reference ``planner.py:403``."""
    src_index, dst_by_rank = index_metadata_rosters(gathered_pairs)
    plan = ReshardPlan([], [])
    for task, key, s, d, s_sl, d_sl, numel in _iter_global_transfer_ops(src_index, dst_by_rank):
        if s.owner_rank == my_global_rank:
            nb = numel * s.element_size
            plan.send_ops.append(RefitTransferOp(s.name, d.owner_rank, True, s_sl, d_sl, task, nb, s.dtype))
            plan.send_bytes[d.owner_rank] = plan.send_bytes.get(d.owner_rank, 0) + nb
        if d.owner_rank == my_global_rank:
            nb = numel * s.element_size
            op = RefitTransferOp(d.name, s.owner_rank, False, d_sl, s_sl, task, nb, s.dtype)
            plan.recv_ops.append(op)
            plan.recv_bytes[s.owner_rank] = plan.recv_bytes.get(s.owner_rank, 0) + nb
    return plan


def build_local_reshard_plan(src_module, dst_module, src_pg_collection=None, dst_pg_collection=None, num_experts=None, group=None,
                             src_rank_offset: int = 0, dst_rank_offset: int = 0) -> ReshardPlan:
    """Every rank gathers every rank's (source, destination) metadata once and derives ITS send / recv lists
    (reference ``planner.py:454``).  Either module may be ``None`` (non-collocated: pure sender / pure receiver / idle)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    # group-local ranks of the two layouts are shifted into this (joint) world by the offsets
    mine = (extract_module_metadata(src_module, rank - src_rank_offset, src_pg_collection, num_experts, src_rank_offset),
            extract_module_metadata(dst_module, rank - dst_rank_offset, dst_pg_collection, num_experts, dst_rank_offset))
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, mine, group=group)
    else:
        gathered = [mine]
    return build_plan_from_rosters(gathered, rank)


def build_centralized_reshard_plan(src_module, dst_module, src_pg_collection=None, dst_pg_collection=None, num_experts=None, group=None,
                                   src_rank_offset: int = 0, dst_rank_offset: int = 0) -> ReshardPlan:
    """Rank 0 plans for everybody and scatters the per-rank plans (reference ``planner.py:503``) — one O(world) gather and
    one scatter instead of an O(world²) all-gather of metadata; pays off beyond a few hundred ranks."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return build_local_reshard_plan(src_module, dst_module, src_pg_collection, dst_pg_collection, num_experts, group, src_rank_offset, dst_rank_offset)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = (extract_module_metadata(src_module, rank - src_rank_offset, src_pg_collection, num_experts, src_rank_offset),
            extract_module_metadata(dst_module, rank - dst_rank_offset, dst_pg_collection, num_experts, dst_rank_offset))
    root = dist.get_global_rank(group, 0) if group is not None else 0
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=root, group=group)
    plans = [build_plan_from_rosters(gathered, r) for r in range(world)] if rank == 0 else None
    out = [None]
    dist.scatter_object_list(out, plans, src=root, group=group)
    return out[0]
