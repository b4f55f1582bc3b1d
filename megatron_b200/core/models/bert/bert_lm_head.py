"""Masked-LM head: dense → activation → LayerNorm; logits come from the tied output layer (reference ``bert_lm_head.py``)."""
import torch

from ...transformer.module import MegatronModule
from ...transformer.torch_norm import FusedNorm


class BertLMHead(MegatronModule):
    def __init__(self, hidden_size: int, config):
        super().__init__(config)
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.dense = torch.nn.Linear(hidden_size, hidden_size, device=dev, dtype=config.params_dtype)
        if config.perform_initialization:
            config.init_method(self.dense.weight)
            self.dense.bias.data.zero_()
        for p in self.dense.parameters():
            setattr(p, "sequence_parallel", config.sequence_parallel)
        self.layer_norm = FusedNorm(config, hidden_size, eps=config.layernorm_epsilon, normalization="LayerNorm")
        self.gelu = torch.nn.functional.gelu

    def forward(self, hidden_states):
        return self.layer_norm(self.gelu(self.dense(hidden_states)))
