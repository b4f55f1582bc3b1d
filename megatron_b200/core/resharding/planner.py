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
