"""Pipeline / tensor-parallel plumbing of the serving path (reference ``inference/communication_utils.py``)."""
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import parallel_state as ps


def is_pipeline_first_stage(pp_group=None) -> bool:
    return ps.is_pipeline_first_stage() if pp_group is None else dist.get_rank(pp_group) == 0


def is_pipeline_last_stage(pp_group=None) -> bool:
    return ps.is_pipeline_last_stage() if pp_group is None else dist.get_rank(pp_group) == dist.get_world_size(pp_group) - 1


def broadcast_from_last_pipeline_stage(size, dtype, tensor: Optional[torch.Tensor] = None, pp_group=None) -> torch.Tensor:
    """Every pipeline stage ends up with the last stage's ``tensor`` (logits / sampled tokens)."""
    pp_group = pp_group or ps.get_pipeline_model_parallel_group()
    if dist.get_world_size(pp_group) == 1:
        return tensor
    if not is_pipeline_last_stage(pp_group):
        tensor = torch.empty(size, dtype=dtype, device="cuda" if dist.get_backend(pp_group) == "nccl" else "cpu")
    dist.broadcast(tensor, src=dist.get_process_group_ranks(pp_group)[-1], group=pp_group)
    return tensor


def send_to_next_pipeline_rank(tensor: torch.Tensor) -> None:
    dist.send(tensor, ps.get_pipeline_model_parallel_next_rank())


def recv_from_prev_pipeline_rank_(recv_buffer: torch.Tensor) -> None:
    dist.recv(recv_buffer, ps.get_pipeline_model_parallel_prev_rank())


def broadcast_tensor(size, dtype, tensor: Optional[torch.Tensor] = None, rank: int = 0, data_parallel: bool = False) -> torch.Tensor:
    """From ``rank`` to the whole world (or, with ``data_parallel``, from the first rank of every model-parallel group to its group)."""
    group = ps.get_model_parallel_group() if data_parallel else None
    src = dist.get_process_group_ranks(group)[0] if data_parallel else rank
    if dist.get_rank() != src:
        tensor = torch.empty(size, dtype=dtype, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    dist.broadcast(tensor, src, group=group)
    return tensor


def _broadcast_list(values: Optional[List], dtype, rank: int, data_parallel: bool) -> List:
    n = torch.tensor([len(values) if values is not None else 0], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    n = broadcast_tensor(1, torch.int64, n, rank, data_parallel)
    t = torch.tensor(values, dtype=dtype, device=n.device) if values is not None else None
    return broadcast_tensor(int(n.item()), dtype, t, rank, data_parallel).tolist()


def broadcast_int_list(size=None, int_list: Optional[List[int]] = None, rank: int = 0, data_parallel: bool = False) -> List[int]:
    return _broadcast_list(int_list, torch.int64, rank, data_parallel)


def broadcast_float_list(size=None, float_list: Optional[List[float]] = None, rank: int = 0, data_parallel: bool = False) -> List[float]:
    return _broadcast_list(float_list, torch.float32, rank, data_parallel)
