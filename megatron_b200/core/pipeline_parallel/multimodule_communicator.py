"""Pipeline communication for models whose modules sit on different rank grids (reference ``pipeline_parallel/multimodule_communicator.py`` —
``MultiModulePipelineCommunicator`` :110, stage counting :506-600).

``module_to_grid_map`` places every module on a ``HyperCommGrid``; ``topology[name]`` lists the modules that consume ``name``'s output
(a DAG, e.g. ``{"vision": ["language"], "audio": ["language"], "language": []}``).  Between two modules a ``BridgeCommunicator`` re-partitions
the batch between their DP layouts; inside a module with ``pp > 1`` plain stage-to-stage point-to-point is used.  The API mirrors
``P2PCommunicator`` but moves dictionaries keyed by module name, because a sink may have several producers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .bridge_communicator import BridgeCommunicator


@dataclass
class RankModuleInfo:
    name: str
    pp_rank: int
    pp_size: int
    is_source: bool                 # nothing feeds this module
    is_sink: bool                   # feeds nothing
    bridges_in: List[str] = field(default_factory=list)
    bridges_out: List[str] = field(default_factory=list)


def _pp_coord(grid, rank: int) -> Tuple[int, int]:
    if "pp" not in grid.dim_names:
        return 0, 1
    st = 1
    for name, n in zip(grid.dim_names, grid.shape):
        if name == "pp":
            return ((rank - grid.rank_offset) // st) % n, n
        st *= n
    return 0, 1


class MultiModulePipelineCommunicator:
    def __init__(self, module_to_grid_map: Dict[str, object], topology: Dict[str, List[str]], config=None, dim_mapping: Optional[Dict[str, int]] = None):
        self.grids, self.topology, self.config = module_to_grid_map, topology, config
        self.rank = dist.get_rank()
        self.dtype = getattr(config, "pipeline_dtype", None) or torch.float32
        producers: Dict[str, List[str]] = {n: [] for n in module_to_grid_map}
        for src, dsts in topology.items():
            for d in dsts:
                producers[d].append(src)
        self.producers = producers
        # bridges are created collectively, in sorted edge order, on every rank
        self.bridges: Dict[Tuple[str, str], BridgeCommunicator] = {}
        for src in sorted(topology):
            for dst in sorted(topology[src]):
                self.bridges[(src, dst)] = BridgeCommunicator(module_to_grid_map[src], module_to_grid_map[dst], dim_mapping=dim_mapping)
        self.my_modules: Dict[str, RankModuleInfo] = {}
        for name, g in module_to_grid_map.items():
            if g.rank_offset <= self.rank < g.rank_offset + g.size:
                pr, ps = _pp_coord(g, self.rank)
                self.my_modules[name] = RankModuleInfo(name, pr, ps, not producers[name], not topology.get(name), producers[name], list(topology.get(name, [])))

    # ---- stage bookkeeping ----
    def is_current_rank_in_grid(self, grid) -> bool:
        return grid.rank_offset <= self.rank < grid.rank_offset + grid.size

    @property
    def is_pp_first_stage(self) -> bool:
        return any(m.is_source and m.pp_rank == 0 for m in self.my_modules.values())

    @property
    def is_pp_last_stage(self) -> bool:
        return any(m.is_sink and m.pp_rank == m.pp_size - 1 for m in self.my_modules.values())

    @staticmethod
    def compute_total_pipeline_stages(topology: Dict[str, List[str]], module_pp: Dict[str, int]) -> int:
        """Length (in pipeline stages) of the longest producer→consumer chain: what the 1F1B warm-up depth must cover."""
        memo: Dict[str, int] = {}

        def longest_from(n: str) -> int:
            if n not in memo:
                memo[n] = module_pp.get(n, 1) + max((longest_from(d) for d in topology.get(n, [])), default=0)
            return memo[n]

        return max(longest_from(n) for n in topology)

    @property
    def total_stages(self) -> int:
        return self.compute_total_pipeline_stages(self.topology, {n: _pp_coord(g, g.rank_offset)[1] for n, g in self.grids.items()})

    # ---- communication ----
    def recv_forward(self, tensor_shapes: Dict[str, Sequence[int]]) -> Dict[str, torch.Tensor]:
        """Inputs of the modules hosted here: ``{producer_name: tensor}`` on a module's first stage."""
        out = {}
        for m in self.my_modules.values():
            if m.pp_rank == 0:
                for src in m.bridges_in:
                    out[src] = self.bridges[(src, m.name)].recv_forward(tensor_shapes[src], self.dtype)
        return out

    def send_forward(self, output_dict: Dict[str, torch.Tensor]) -> None:
        for m in self.my_modules.values():
            if m.pp_rank == m.pp_size - 1 and m.name in output_dict:
                for dst in m.bridges_out:
                    self.bridges[(m.name, dst)].send_forward(output_dict[m.name])

    def recv_backward(self, tensor_shapes: Dict[str, Sequence[int]]) -> Dict[str, torch.Tensor]:
        out = {}
        for m in self.my_modules.values():
            if m.pp_rank == m.pp_size - 1:
                for dst in m.bridges_out:
                    g = self.bridges[(m.name, dst)].recv_backward(tensor_shapes[m.name], self.dtype)
                    out[m.name] = g if m.name not in out else out[m.name] + g       # several consumers: gradients add
        return out

    def send_backward(self, grad_dict: Dict[str, torch.Tensor]) -> None:
        for m in self.my_modules.values():
            if m.pp_rank == 0:
                for src in m.bridges_in:
                    if src in grad_dict:
                        self.bridges[(src, m.name)].send_backward(grad_dict[src])

    def send_forward_recv_backward(self, output_dict, tensor_shapes):
        self.send_forward(output_dict)
        return self.recv_backward(tensor_shapes)

    def send_backward_recv_forward(self, grad_dict, tensor_shapes):
        self.send_backward(grad_dict)
        return self.recv_forward(tensor_shapes)
