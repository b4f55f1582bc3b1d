"""Hybrid layer pattern: ``M`` = Mamba, ``G`` = gated delta net, ``*`` = attention, ``-`` = MLP, ``E`` = MoE (reference ``models/hybrid/hybrid_layer_allocation.py``
``Symbols``); ``+`` = multi-latent attention, ``D`` = DeepSeek sparse attention; ``|`` marks pipeline-stage boundaries (``"M*M*|M*M-"``: the first four
layers on stage 0, the rest on stage 1 — uneven splits allowed) and everything after ``/`` describes multi-token-prediction layers."""
from __future__ import annotations

from typing import List, Optional


class Symbols:
    MAMBA, ATTENTION, MLP, MOE, GDN, MLA, DS_ATTENTION = "M", "*", "-", "E", "G", "+", "D"
    PIPE, MTP_SEPARATOR = "|", "/"
    VALID = {MAMBA, ATTENTION, MLP, MOE, GDN, MLA, DS_ATTENTION}


def parse_hybrid_pattern(pattern: str):
    """``"M*|M-/MM"`` → (flat main layout ``['M','*','M','-']``, layers per pipeline segment ``[2, 2]`` or None when there is no ``|``, MTP patterns ``['MM']``)."""
    main, *mtp = pattern.split(Symbols.MTP_SEPARATOR)
    segments = main.split(Symbols.PIPE)
    return list(main.replace(Symbols.PIPE, "")), ([len(s) for s in segments] if len(segments) > 1 else None), mtp


def _allocate_auto(total: int, attn_ratio: float, mlp_ratio: float) -> List[str]:
    n_attn, n_mlp = round(total * attn_ratio), round(total * mlp_ratio)
    n_mamba = total - n_attn - n_mlp
    assert n_mamba >= 0
    layout = [Symbols.MAMBA] * total
    # spread attention layers evenly, then MLP layers evenly over the remaining slots
    if n_attn:
        step = total / n_attn
        for i in range(n_attn):
            layout[min(total - 1, int(step * i + step / 2))] = Symbols.ATTENTION
    free = [i for i, s in enumerate(layout) if s == Symbols.MAMBA]
    if n_mlp:
        step = len(free) / n_mlp
        for i in range(n_mlp):
            layout[free[min(len(free) - 1, int(step * i + step / 2))]] = Symbols.MLP
    return layout


def allocate_layers(total_layers: int, target_attention_ratio: float = 0.0, target_mlp_ratio: float = 0.0, override_pattern: Optional[str] = None) -> List[str]:
    if override_pattern:
        layout = parse_hybrid_pattern(override_pattern)[0]
        bad = set(layout) - Symbols.VALID
        if bad:
            raise ValueError(f"invalid symbols {bad} in hybrid pattern; valid: {sorted(Symbols.VALID)}")
        if len(layout) != total_layers:
            raise ValueError(f"hybrid pattern has {len(layout)} layers, model has {total_layers}")
        return layout
    return _allocate_auto(total_layers, target_attention_ratio, target_mlp_ratio)
