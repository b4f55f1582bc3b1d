"""Configuration of a multi-in / multi-out model (reference ``models/mimo/config/base_configs.py``, ``config/role.py``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

from ...transformer.spec_utils import ModuleSpec

MIMO_LANGUAGE_MODULE_KEY = "language"


@dataclass
class MimoModelConfig:
    """``language_model_spec`` builds the decoder; ``modality_submodules_spec[name]`` builds a ``ModalitySubmodules`` (encoders +
    projections); ``special_token_ids[name]`` is the placeholder id in ``input_ids`` where that modality's embeddings go.
    ``module_to_grid_map`` (optional) places every module on its own ``HyperCommGrid``: identical rank sets → colocated, disjoint
    rank sets → the modules form pipeline stages connected by ``BridgeCommunicator``s."""

    language_model_spec: ModuleSpec = field(default_factory=lambda: ModuleSpec(module=None))
    modality_submodules_spec: Dict[str, ModuleSpec] = field(default_factory=dict)
    special_token_ids: Dict[str, int] = field(default_factory=dict)
    module_to_grid_map: Optional[Dict[str, object]] = None
    kv_format: str = "sbhd"

    def __post_init__(self):
        if self.module_to_grid_map:
            want = set(self.modality_submodules_spec) | {MIMO_LANGUAGE_MODULE_KEY}
            have = set(self.module_to_grid_map)
            if want != have:
                raise ValueError(f"module_to_grid_map keys must be the modality names + '{MIMO_LANGUAGE_MODULE_KEY}': missing {want - have}, extra {have - want}")
        missing = set(self.modality_submodules_spec) - set(self.special_token_ids)
        if missing:
            raise ValueError(f"no special token id for modalities {sorted(missing)}")

    def role_of(self, rank: int) -> Dict[str, bool]:
        """Which modules this rank hosts (all of them when no grids are given)."""
        names = list(self.modality_submodules_spec) + [MIMO_LANGUAGE_MODULE_KEY]
        if not self.module_to_grid_map:
            return {n: True for n in names}
        return {n: g.rank_offset <= rank < g.rank_offset + g.size for n, g in self.module_to_grid_map.items()}
