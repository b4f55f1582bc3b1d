"""Common DDP interface (reference ``distributed/data_parallel_base.py``)."""
from contextlib import contextmanager

import torch

from ..transformer.module import MegatronModule


class _BaseDataParallel(MegatronModule):
    def __init__(self, config, module: torch.nn.Module):
        super().__init__(config=config)
        self.module = module

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    @contextmanager
    def no_sync(self):
        try:
            yield
        finally:
            pass

    def start_grad_sync(self, *unused):
        pass

    def scale_gradients(self, scaling_factor: float) -> None:
        pass

    def finish_grad_sync(self, force_all_reduce=False):
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
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def sharded_state_dict(self, prefix="", *args, **kwargs):
        return self.module.sharded_state_dict(prefix, *args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        self.module.load_state_dict(state_dict, strict=strict)

    def set_input_tensor(self, t):
        return self.module.set_input_tensor(t)
