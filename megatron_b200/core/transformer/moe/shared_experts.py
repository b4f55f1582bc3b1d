"""Always-on shared expert (DeepSeek / Qwen-MoE style; reference ``moe/shared_experts.py``)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..mlp import MLP, MLPSubmodules
from ..transformer_config import TransformerConfig


class SharedExpertMLP(MLP):
    """Dense MLP applied to every token, optionally gated by ``sigmoid(x Wg)``.  With
    ``moe_shared_expert_overlap`` it runs on a side stream while the routed tokens are in flight."""

    stream = None

    def __init__(self, config: TransformerConfig, submodules: MLPSubmodules, gate: bool = False, pg_collection=None):
        super().__init__(config, submodules, ffn_hidden_size=config.moe_shared_expert_intermediate_size,
                         tp_group=getattr(pg_collection, "tp", None) if pg_collection is not None else None)
        self.use_shared_expert_gate = gate or config.moe_shared_expert_gate
        if self.use_shared_expert_gate:
            dev = self.linear_fc1.weight.device
            self.gate_weight = torch.nn.Parameter(torch.empty((1, config.hidden_size), device=dev, dtype=config.params_dtype))
            if config.perform_initialization:
                config.init_method(self.gate_weight)
            setattr(self.gate_weight, "sequence_parallel", config.sequence_parallel)
        else:
            self.gate_weight = None
        self._pending = None

    def forward(self, hidden_states):
        out, _ = super().forward(hidden_states)
        if self.use_shared_expert_gate:
            out = out * torch.sigmoid(F.linear(hidden_states, self.gate_weight.to(hidden_states.dtype)))
        return out

    # overlap API: launch on a side stream, join after the routed experts
    def launch(self, hidden_states):
        if not (self.config.moe_shared_expert_overlap and hidden_states.is_cuda):
            self._pending = ("sync", self.forward(hidden_states))
            return
        if SharedExpertMLP.stream is None:
            SharedExpertMLP.stream = torch.cuda.Stream()
        s = SharedExpertMLP.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = self.forward(hidden_states)
        self._pending = ("async", out)

    def join(self):
        kind, out = self._pending
        self._pending = None
        if kind == "async":
            torch.cuda.current_stream().wait_stream(SharedExpertMLP.stream)
            out.record_stream(torch.cuda.current_stream())
        return out
