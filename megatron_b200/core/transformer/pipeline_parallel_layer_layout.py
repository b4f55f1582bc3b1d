"""Explicit per-stage layer layouts for pipeline parallelism (reference ``transformer/pipeline_parallel_layer_layout.py`` 321 LoC).

A layout string lists the contents of every (pp stage × virtual chunk) separated by ``|``:
``E`` embedding, ``t`` decoder layer, ``m`` MTP layer, ``L`` loss/output head; ``x*3`` repeats, parentheses group:
``"Et*3|(tt|)*29,m|L"`` → stage 0: embedding + 3 layers, then 29 stages of 2 layers, one stage with an MTP layer, a last stage with the
loss.  Stages are assigned round-robin to virtual chunks: stage index ``s`` lives on pp rank ``s % pp`` as virtual chunk ``s // pp``."""
from __future__ import annotations

import re
from typing import List, Optional

from ..enums import LayerType

_SYMBOL = {"E": LayerType.embedding, "t": LayerType.decoder, "m": LayerType.mtp, "L": LayerType.loss}


def _expand(s: str) -> str:
    """Resolve ``(...)*n`` groups and ``x*n`` repeats into a flat string."""
    s = s.replace(",", "")
    pat_group = re.compile(r"\(([^()]*)\)\*(\d+)")
    while True:
        m = pat_group.search(s)
        if not m:
            break
        s = s[: m.start()] + m.group(1) * int(m.group(2)) + s[m.end():]
    s = s.replace("(", "").replace(")", "")
    pat_rep = re.compile(r"([EtmL])\*(\d+)")
    while True:
        m = pat_rep.search(s)
        if not m:
            break
        s = s[: m.start()] + m.group(1) * int(m.group(2)) + s[m.end():]
    return s


class PipelineParallelLayerLayout:
    def __init__(self, layout: List[List[LayerType]], pipeline_model_parallel_size: int):
        self.flat = layout
        self.pp = pipeline_model_parallel_size
        assert len(layout) % self.pp == 0, f"{len(layout)} stages cannot be divided over pp={self.pp}"
        self.vpp = len(layout) // self.pp
        # layout[pp_rank][vp_stage]
        self.layout = [[layout[v * self.pp + r] for v in range(self.vpp)] for r in range(self.pp)]

    @classmethod
    def from_str(cls, layout: str, pipeline_model_parallel_size: int) -> "PipelineParallelLayerLayout":
        stages = _expand(layout).split("|")
        if stages and stages[-1] == "" and layout.rstrip().endswith("|"):
            stages = stages[:-1]
        parsed = []
        for st in stages:
            bad = set(st) - set(_SYMBOL)
            if bad:
                raise ValueError(f"unknown layout symbols {bad} in stage '{st}'")
            parsed.append([_SYMBOL[c] for c in st])
        return cls(parsed, pipeline_model_parallel_size)

    def validate_layer_layout(self, num_layers: int, mtp_num_layers: Optional[int] = 0):
        flat = [x for st in self.flat for x in st]
        assert flat.count(LayerType.embedding) == 1 and LayerType.embedding in self.flat[0], "the embedding must be on the first stage"
        assert flat.count(LayerType.loss) == 1 and LayerType.loss in self.flat[-1], "the loss must be on the last stage"
        assert flat.count(LayerType.decoder) == num_layers, f"layout has {flat.count(LayerType.decoder)} decoder layers, config has {num_layers}"
        assert flat.count(LayerType.mtp) == (mtp_num_layers or 0), "MTP layer count mismatch"
        # MTP layers consume the final hidden state: none may precede the last decoder layer, and (standalone placement) a stage made only of
        # ``m`` symbols is legal as long as it sits between the last decoder layer and the loss
        order = [x for x in flat if x in (LayerType.decoder, LayerType.mtp)]
        if LayerType.mtp in order:
            first_m = order.index(LayerType.mtp)
            assert LayerType.decoder not in order[first_m:], "MTP layers must come after every decoder layer"

    def mtp_standalone_stages(self) -> List[int]:
        """Flat stage indices that hold MTP layers and nothing else (reference ``mtp_standalone``)."""
        return [i for i, st in enumerate(self.flat) if st and all(x == LayerType.mtp for x in st)]

    def get_num_layers_to_build(self, layer_type: LayerType = LayerType.decoder, vp_stage: Optional[int] = None, pp_rank: int = 0) -> int:
        return self.layout[pp_rank][vp_stage or 0].count(layer_type)

    def get_layer_offset(self, layer_type: LayerType = LayerType.decoder, vp_stage: Optional[int] = None, pp_rank: int = 0) -> int:
        idx = (vp_stage or 0) * self.pp + pp_rank
        return sum(st.count(layer_type) for st in self.flat[:idx])

    def get_layer_id_list(self, layer_type: LayerType = LayerType.decoder, vp_stage: Optional[int] = None, pp_rank: int = 0) -> List[int]:
        off = self.get_layer_offset(layer_type, vp_stage, pp_rank)
        return list(range(off, off + self.get_num_layers_to_build(layer_type, vp_stage, pp_rank)))

    def pretty_repr(self) -> str:
        inv = {v: k for k, v in _SYMBOL.items()}
        return "\\n".join(f"pp{r}: " + " | ".join("".join(inv[x] for x in st) or "-" for st in self.layout[r]) for r in range(self.pp))
