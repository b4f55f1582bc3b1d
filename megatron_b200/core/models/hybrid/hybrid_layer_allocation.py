"""Unified hybrid layer pattern: ``"M-M-|M-M*-/MM/MM"`` = main decoder (two pipeline segments) + two MTP depths of pattern ``MM``
(reference ``models/hybrid/hybrid_layer_allocation.py:14-530``).  The symbol table and the flat parser live with the Mamba
stack (``core/ssm/mamba_hybrid_layer_allocation.py``); this module adds the structured view the model builders use."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from ...ssm.mamba_hybrid_layer_allocation import Symbols, allocate_layers  # noqa: F401
from ...ssm.mamba_hybrid_layer_allocation import parse_hybrid_pattern as parse_flat_hybrid_pattern  # noqa: F401


@dataclass
class ParsedHybridPattern:
    main_pattern: Optional[str]          # may contain '|'
    mtp_pattern: Optional[str]           # pattern of ONE MTP depth (all depths are identical)
    mtp_num_depths: int

    @property
    def layer_types(self) -> List[str]:
        return [c for c in (self.main_pattern or "") if c != Symbols.PIPE]

    @property
    def segments(self) -> List[str]:
        return (self.main_pattern or "").split(Symbols.PIPE)


def _validate_pattern(pattern: str, name: str, allow_pipe: bool = False) -> None:
    ok = set(Symbols.VALID) | ({Symbols.PIPE} if allow_pipe else set())
    bad = sorted(set(pattern) - ok)
    if bad:
        raise ValueError(f"{name} '{pattern}' contains invalid symbol(s) {bad}; valid: {sorted(ok)}")


def parse_hybrid_pattern(pattern: Optional[str]) -> ParsedHybridPattern:
    """Split at '/': the first part is the decoder, every further part one MTP depth (they must all be equal)."""
    if pattern is None:
        return ParsedHybridPattern(None, None, 0)
    main, *mtp = pattern.split(Symbols.MTP_SEPARATOR)
    _validate_pattern(main, "main pattern", allow_pipe=True)
    if not mtp:
        return ParsedHybridPattern(main, None, 0)
    if any(m != mtp[0] for m in mtp):
        raise ValueError(f"all MTP depths must use the same pattern, got {mtp}")
    if not mtp[0]:
        raise ValueError("empty MTP pattern after '/'")
    _validate_pattern(mtp[0], "MTP pattern")
    return ParsedHybridPattern(main, mtp[0], len(mtp))


def pattern_from_ratios(num_layers: int, attention_ratio: float = 0.0, mlp_ratio: float = 0.0) -> str:
    """The deprecated ``--hybrid-attention-ratio`` / ``--hybrid-mlp-ratio`` arguments as a pattern string.  Checkpoints trained
    with the ratios depend on the exact placement, so this reproduces the reference's rule: walk the layers with an accumulator
    that drops by one per Mamba layer and is topped up by the average gap whenever it falls below one half — the stack
    starts and ends with Mamba layers and the special layers are (nearly) evenly spaced; MLP layers are then placed the
    same way among the remaining Mamba slots."""
    assert num_layers > 0 and 0.0 <= attention_ratio <= 1.0 and 0.0 <= mlp_ratio <= 1.0 and attention_ratio + mlp_ratio <= 1.0
    types = [Symbols.MAMBA] * num_layers

    def place(symbol: str, gap: float, eligible) -> None:
        acc = gap
        for i in range(num_layers):
            if not eligible(i):
                continue
            if acc < 0.5:
                types[i] = symbol
                acc += gap
            else:
                acc -= 1

    n_attn = round(num_layers * attention_ratio)
    n_mamba = num_layers - n_attn
    place(Symbols.ATTENTION, n_mamba / (n_attn + 1), lambda i: True)
    n_mlp = round(num_layers * mlp_ratio)
    if n_mlp > 0:
        place(Symbols.MLP, (n_mamba - n_mlp) / n_mlp, lambda i: types[i] == Symbols.MAMBA)
    return "".join(types)


def get_hybrid_total_layer_count(pattern: str) -> int:
    """Decoder layers of a unified pattern (pipes and MTP parts do not count)."""
    return len(parse_hybrid_pattern(pattern).layer_types)


def get_hybrid_total_pipeline_segment_count(pattern: str) -> int:
    return len(parse_hybrid_pattern(pattern).segments)


def get_hybrid_layer_counts(pattern: str) -> Dict[str, int]:
    """Layers per symbol over the decoder AND all MTP depths (what the FLOPs / memory models need)."""
    p = parse_hybrid_pattern(pattern)
    counts = {s: 0 for s in sorted(Symbols.VALID)}
    for c in p.layer_types:
        counts[c] += 1
    for c in (p.mtp_pattern or ""):
        counts[c] += p.mtp_num_depths
    return counts


def validate_segment_layers(segment: str) -> List[str]:
    _validate_pattern(segment, "pipeline segment")
    return list(segment)


def select_pipeline_segment(main_pattern: str, pp_group=None, vp_stage: Optional[int] = None, first_stage_layers: Optional[int] = None,
                            last_stage_layers: Optional[int] = None, tp_group=None, dp_cp_group=None, *, pp_rank: Optional[int] = None,
                            pp_size: Optional[int] = None) -> Tuple[List[str], int]:
    """(layer symbols of THIS pipeline rank / virtual stage, number of layers before them).

    With '|' in the pattern the segments are taken as written; segment ``vp_stage * pp_size + pp_rank`` belongs to this rank
    (interleaved schedule order) and the number of segments must be a multiple of the pipeline size.  Without '|' the layers
    are sliced evenly, or unevenly with ``first_stage_layers`` / ``last_stage_layers``."""
    import torch.distributed as dist
    if pp_size is None:
        pp_size = dist.get_world_size(pp_group) if (pp_group is not None and dist.is_initialized()) else 1
    if pp_rank is None:
        pp_rank = dist.get_rank(pp_group) if (pp_group is not None and dist.is_initialized()) else 0
    segs = main_pattern.split(Symbols.PIPE)
    if len(segs) > 1:
        if first_stage_layers is not None or last_stage_layers is not None:
            raise ValueError("first/last stage layer counts cannot be combined with '|' stage boundaries")
        if len(segs) % pp_size:
            raise ValueError(f"{len(segs)} pipeline segments do not divide over {pp_size} pipeline ranks")
        vp = len(segs) // pp_size
        if vp > 1 and vp_stage is None:
            raise ValueError(f"the pattern defines {vp} virtual stages per rank: vp_stage is required")
        idx = (vp_stage or 0) * pp_size + pp_rank
        return validate_segment_layers(segs[idx]), sum(len(s) for s in segs[:idx])
    layers = validate_segment_layers(main_pattern)
    n = len(layers)
    if pp_size == 1:
        return layers, 0
    if vp_stage not in (None, 0) and (first_stage_layers is not None or last_stage_layers is not None):
        raise ValueError("uneven first/last stages with virtual pipeline stages need explicit '|' boundaries")
    if first_stage_layers is None and last_stage_layers is None:
        vp = 1
        if n % pp_size:
            raise ValueError(f"{n} layers do not divide evenly over {pp_size} stages: mark the boundaries with '|' or give first/last stage sizes")
        per = n // pp_size
        return layers[pp_rank * per:(pp_rank + 1) * per], pp_rank * per
    first = first_stage_layers if first_stage_layers is not None else None
    last = last_stage_layers if last_stage_layers is not None else None
    middle_ranks = pp_size - (first is not None) - (last is not None)
    rest = n - (first or 0) - (last or 0)
    if middle_ranks and rest % middle_ranks:
        raise ValueError(f"{rest} remaining layers do not divide over {middle_ranks} middle stages")
    per = rest // middle_ranks if middle_ranks else 0
    sizes = [(first if (r == 0 and first is not None) else last if (r == pp_size - 1 and last is not None) else per) for r in range(pp_size)]
    if sum(sizes) != n:
        raise ValueError(f"stage sizes {sizes} do not add up to {n} layers")
    off = sum(sizes[:pp_rank])
    return layers[off:off + sizes[pp_rank]], off


def get_layer_maps_from_layer_type_list(layer_type_list: List[str]) -> Dict[str, Dict[int, int]]:
    """Per symbol: global layer index -> index among the layers of that type (e.g. which KV-cache slot attention layer 7 uses,
    which recurrent-state slot Mamba layer 3 uses)."""
    maps: Dict[str, Dict[int, int]] = {s: {} for s in sorted(Symbols.VALID)}
    for g, t in enumerate(layer_type_list):
        maps[t][g] = len(maps[t])
    return maps
