"""Adapter around PyTorch FSDP2 (``torch.distributed.fsdp.fully_shard``) — reference ``distributed/torch_fully_sharded_data_parallel.py``.

Use when DTensor-based sharding is wanted (interop with torch.distributed.checkpoint, torch-native tooling); the in-house ``fsdp/`` package is the
default because it composes with this framework's fp32-master optimizers.  Every ``TransformerLayer`` (configurable) becomes an FSDP unit, the root
holds the rest; gradient reduction happens inside FSDP2, so the DDP-style hooks are no-ops."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .data_parallel_base import _BaseDataParallel


class TorchFullyShardedDataParallel(_BaseDataParallel):
    def __init__(self, config, ddp_config, module: torch.nn.Module, sub_modules_to_wrap: Optional[Sequence[type]] = None, process_group=None, mesh=None,
                 reshard_after_forward: bool = True, **_):
        super().__init__(config, module)
        from torch.distributed.fsdp import fully_shard

        if sub_modules_to_wrap is None:
            from ..transformer.transformer_layer import TransformerLayer

            sub_modules_to_wrap = (TransformerLayer,)
        if mesh is None:
            from torch.distributed.device_mesh import DeviceMesh

            from .. import parallel_state as ps

            group = process_group if process_group is not None else (ps.get_data_parallel_group(with_context_parallel=True) if ps.is_initialized() else None)
            ranks = torch.distributed.get_process_group_ranks(group) if group is not None else list(range(torch.distributed.get_world_size()))
            dev = "cuda" if torch.cuda.is_available() and torch.distributed.get_backend(group) == "nccl" else "cpu"
            mesh = DeviceMesh(dev, ranks)
        self.ddp_config, self.mesh = ddp_config, mesh
        kw = dict(mesh=mesh, reshard_after_forward=reshard_after_forward)
        for sub in module.modules():
            if isinstance(sub, tuple(sub_modules_to_wrap)):
                fully_shard(sub, **kw)
        fully_shard(self.module, **kw)

    def load_state_dict(self, state_dict, strict=True):
        from torch.distributed.checkpoint.state_dict import StateDictOptions, set_model_state_dict

        return set_model_state_dict(self.module, state_dict, options=StateDictOptions(full_state_dict=True, strict=strict))

    def full_state_dict(self):
        from torch.distributed.checkpoint.state_dict import StateDictOptions, get_model_state_dict

        return get_model_state_dict(self.module, options=StateDictOptions(full_state_dict=True))
