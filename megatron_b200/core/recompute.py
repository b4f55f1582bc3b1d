"""Full-layer activation recompute helpers (reference ``core/recompute.py:21-189``): ``uniform`` and ``block`` methods over a list of
layers, built on the RNG-state-preserving ``tensor_parallel.checkpoint``."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import tensor_parallel


def checkpointed_forward(layers: List[torch.nn.Module], hidden_states: torch.Tensor, *, method: str = "uniform", num_layers: Optional[int] = None,
                         layer_forward: Optional[Callable] = None, distribute_saved_activations: bool = False, **kwargs) -> torch.Tensor:
    """Run ``layers`` on ``hidden_states`` recomputing activations in backward.

    ``uniform``: checkpoint every group of ``num_layers`` consecutive layers (default 1).
    ``block``  : checkpoint only the first ``num_layers`` layers, run the rest normally."""
    n = len(layers)
    num_layers = num_layers or 1

    def run(start, end):
        def fwd(h):
            for i in range(start, end):
                out = layer_forward(layers[i], h, **kwargs) if layer_forward is not None else layers[i](h, **kwargs)
                h = out[0] if isinstance(out, tuple) else out
            return h

        return fwd

    if method == "uniform":
        i = 0
        while i < n:
            hidden_states = tensor_parallel.checkpoint(run(i, min(n, i + num_layers)), distribute_saved_activations, hidden_states)
            i += num_layers
    elif method == "block":
        for i in range(n):
            if i < num_layers:
                hidden_states = tensor_parallel.checkpoint(run(i, i + 1), distribute_saved_activations, hidden_states)
            else:
                hidden_states = run(i, i + 1)(hidden_states)
    else:
        raise ValueError(f"invalid recompute method {method}")
    return hidden_states
