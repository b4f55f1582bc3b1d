"""Helpers of Megatron-FSDP (reference ``megatron_fsdp/utils.py:1-865``): the distributed index (which process group plays which
role), TP-attribute probes that work without importing the rest of the framework, a name-keyed scratch-memory pool."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class FSDPDistributedIndex:
    """Roles of the process groups around one FSDP instance.

    ``dp_shard`` — parameters / gradients / optimizer state are sharded over it (inside one NVLink domain on B200 systems);
    ``dp_outer`` — HSDP replication (or, with ``hsdp_outer_dp_shard``, a second level of optimizer-state sharding) across domains;
    ``tp`` — tensor parallel group the module was built with;  ``expt_dp_shard`` — the sharding group of expert parameters.
    Built from explicit groups or from a ``torch.distributed.DeviceMesh`` + dimension names."""

    def __init__(self, device_mesh=None, dp_shard_dim: Optional[str] = None, dp_outer_dim: Optional[str] = None, tp_dim: Optional[str] = None,
                 hybrid_fsdp_group=None, hsdp_outer_dp_shard: bool = False, expt_device_mesh=None, dp_shard_group=None, dp_outer_group=None, tp_group=None,
                 expt_dp_shard_group=None):
        self.device_mesh, self.expt_device_mesh = device_mesh, expt_device_mesh
        self.dp_shard_dim, self.dp_outer_dim, self.tp_dim = dp_shard_dim, dp_outer_dim, tp_dim
        self.hsdp_outer_dp_shard = hsdp_outer_dp_shard
        g = lambda mesh, dim: mesh.get_group(dim) if (mesh is not None and dim is not None) else None   # noqa: E731
        self.fsdp_group = dp_shard_group if dp_shard_group is not None else g(device_mesh, dp_shard_dim)
        self.outer_fsdp_group = dp_outer_group if dp_outer_group is not None else g(device_mesh, dp_outer_dim)
        self.tp_group = tp_group if tp_group is not None else g(device_mesh, tp_dim)
        self.expt_fsdp_group = expt_dp_shard_group if expt_dp_shard_group is not None else g(expt_device_mesh, dp_shard_dim)
        self.hybrid_fsdp_group = hybrid_fsdp_group
        if dp_outer_dim is not None and device_mesh is not None and dp_shard_dim is not None:
            assert contains_submesh(device_mesh, (dp_outer_dim, dp_shard_dim)), f"mesh {get_mesh_names(device_mesh)} lacks ({dp_outer_dim}, {dp_shard_dim})"
        self.use_hybrid_fsdp = self.outer_fsdp_group is not None

    def get_fsdp_group(self, is_expert_parallel: bool = False):
        return self.expt_fsdp_group if (is_expert_parallel and self.expt_fsdp_group is not None) else self.fsdp_group

    def get_outer_fsdp_group(self):
        return self.outer_fsdp_group

    def get_dp_group(self, is_expert_parallel: bool = False):
        """The full data-parallel group (outer × shard) when one was given, else the shard group."""
        return self.hybrid_fsdp_group if (self.use_hybrid_fsdp and self.hybrid_fsdp_group is not None) else self.get_fsdp_group(is_expert_parallel)

    def get_root_mesh(self, is_expert_parallel: bool = False):
        return self.expt_device_mesh if (is_expert_parallel and self.expt_device_mesh is not None) else self.device_mesh

    def get_logical_hybrid_fsdp_rank(self) -> int:
        """Rank in the (outer, shard) row-major order — the order optimizer-state shards are laid out under HSDP sharding."""
        s = dist.get_rank(self.fsdp_group) if self.fsdp_group is not None else 0
        o = dist.get_rank(self.outer_fsdp_group) if self.outer_fsdp_group is not None else 0
        n = dist.get_world_size(self.fsdp_group) if self.fsdp_group is not None else 1
        return o * n + s


def get_mesh_names(device_mesh) -> Tuple[str, ...]:
    return tuple(device_mesh.mesh_dim_names or ()) if device_mesh is not None else ()


def contains_submesh(device_mesh, dims: Sequence[str]) -> bool:
    names = get_mesh_names(device_mesh)
    return all(d in names for d in dims)


def is_mcore_tensor_model_parallel(param: torch.Tensor) -> bool:
    return bool(getattr(param, "tensor_model_parallel", False))


def get_mcore_tensor_parallel_partition_dim(param: torch.Tensor) -> Optional[int]:
    return int(getattr(param, "partition_dim")) if is_mcore_tensor_model_parallel(param) else None


def is_mcore_tensor_parallel_duplicated(param: torch.Tensor) -> bool:
    """Replicated across TP ranks (norm weights, row-parallel biases): only TP rank 0 counts them in norms / checkpoints."""
    return not is_mcore_tensor_model_parallel(param)


def find_megatron_fsdp(module: torch.nn.Module):
    from .megatron_fsdp import MegatronFSDP
    for m in module.modules():
        if isinstance(m, MegatronFSDP):
            return m
    return None


class GlobalMemoryBuffer:
    """Name-keyed scratch tensors that only grow (same contract as ``core/utils.py::GlobalMemoryBuffer``; duplicated so that the
    FSDP package has no import-time dependency on the rest of the framework)."""

    def __init__(self):
        self.buffer: Dict[Tuple[str, torch.dtype], torch.Tensor] = {}

    def get_tensor(self, shape, dtype: torch.dtype, name: str, device=None) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        key = (name, dtype)
        t = self.buffer.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=dtype, device=device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
            self.buffer[key] = t
        return t[:n].view(*shape)


_GLOBAL_MEMORY_BUFFER: Optional[GlobalMemoryBuffer] = None


def get_global_memory_buffer() -> GlobalMemoryBuffer:
    global _GLOBAL_MEMORY_BUFFER
    if _GLOBAL_MEMORY_BUFFER is None:
        _GLOBAL_MEMORY_BUFFER = GlobalMemoryBuffer()
    return _GLOBAL_MEMORY_BUFFER


def get_cuda_rng_tracker():
    from .....tensor_parallel.random import get_cuda_rng_tracker as g
    return g()


def initialize_rng_tracker(*a, **k):
    from .....tensor_parallel.random import initialize_rng_tracker as f
    return f(*a, **k)
