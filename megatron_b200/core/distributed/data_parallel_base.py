"""Common surface of the data-parallel wrappers (reference ``distributed/data_parallel_base.py``): DDP, FSDP and the torch-FSDP2 adapter all expose
the hooks the training loop and ``finalize_model_grads`` call."""
from __future__ import annotations

from contextlib import contextmanager

import torch


class _BaseDataParallel(torch.nn.Module):
    def __init__(self, config, module: torch.nn.Module):
        super().__init__()
        self.config, self.module = config, module

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    @contextmanager
    def no_sync(self):
        yield

    def start_grad_sync(self, *unused):
        pass

    def finish_grad_sync(self):
        pass

    def scale_gradients(self, scaling_factor: float):
        pass

    def zero_grad_buffer(self):
        pass

    def broadcast_params(self):
        pass

    def start_param_sync(self, *unused, force_sync: bool = False, force_dispatch: bool = False):
        pass

    def state_dict(self, prefix="", keep_vars=False, destination=None):
        return self.module.state_dict(prefix=prefix, keep_vars=keep_vars, destination=destination)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict(prefix=prefix, keep_vars=keep_vars)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)
