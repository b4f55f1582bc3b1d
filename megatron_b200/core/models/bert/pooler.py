"""Pool the hidden state of one token (``[CLS]``) through dense+tanh (reference ``models/bert/pooler.py``)."""
import torch

from ... import tensor_parallel
from ...transformer.module import MegatronModule


class Pooler(MegatronModule):
    def __init__(self, hidden_size: int, init_method, config, sequence_parallel: bool = False):
        super().__init__(config)
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.dense = torch.nn.Linear(hidden_size, hidden_size, device=dev, dtype=config.params_dtype)
        if config.perform_initialization:
            init_method(self.dense.weight)
            self.dense.bias.data.zero_()
        self.sequence_parallel = sequence_parallel

    def forward(self, hidden_states, sequence_index: int = 0):
        # hidden_states [s, b, h]; under SP gather the sequence first
        if self.sequence_parallel:
            hidden_states = tensor_parallel.gather_from_sequence_parallel_region(hidden_states, tensor_parallel_output_grad=False)
        return torch.tanh(self.dense(hidden_states[sequence_index]))
