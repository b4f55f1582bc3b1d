"""Registry of the non-Adam optimizers (reference ``optimizer/emerging_optimizers.py``).  ``OptimizerConfig.optimizer`` selects one; the update rules
live in ``MegatronOptimizer._apply_update`` (fp32 masters, loss scaling, clipping and checkpointing are shared with Adam)."""
from __future__ import annotations

from typing import Dict

EMERGING_OPTIMIZERS: Dict[str, dict] = {
    "muon": {"module": "megatron_b200.core.optimizer.muon", "needs_whole_matrices": True,
             "doc": "orthogonalised momentum (Newton-Schulz) for 2-D hidden weights, AdamW for the rest; pair with LayerWiseDistributedOptimizer under DP"},
    "soap": {"module": "megatron_b200.core.optimizer.soap", "needs_whole_matrices": True,
             "doc": "AdamW in the eigenbasis of Shampoo's Kronecker factors (one-sided above soap_max_precond_dim); AdamW for 1-D / embedding parameters"},
    "lion": {"module": "megatron_b200.core.optimizer.optimizer", "needs_whole_matrices": False, "doc": "sign of interpolated momentum, one state per parameter"},
}


def is_emerging_optimizer(name: str) -> bool:
    return name in EMERGING_OPTIMIZERS


def needs_whole_matrices(name: str) -> bool:
    """True when the update rule is not elementwise, i.e. incompatible with ZeRO-1 flat-range sharding."""
    return EMERGING_OPTIMIZERS.get(name, {}).get("needs_whole_matrices", False)
