"""Runtime services: resharding planner/executor, communication-memory shims, CUDA-graph manager fallback (CPU, gloo)."""
import torch

from dist_utils import run_distributed


def _reshard(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor
    from megatron_b200.core.resharding import reshard_state_dict

    full_w = torch.arange(8 * 6, dtype=torch.float32).view(8, 6)
    full_b = torch.arange(8, dtype=torch.float32)
    # source: sharded along dim 0 (column-parallel), bias replicated;  destination: sharded along dim 1 (row-parallel), bias sharded
    src = {"w": ShardedTensor.from_rank_offsets("w", full_w.chunk(world, 0)[rank].clone(), (0, rank, world)),
           "b": ShardedTensor.from_rank_offsets("b", full_b.clone())}
    dst = {"w": ShardedTensor.from_rank_offsets("w", torch.zeros(8, 6 // world), (1, rank, world)),
           "b": ShardedTensor.from_rank_offsets("b", torch.zeros(8 // world), (0, rank, world))}
    reshard_state_dict(src, dst, dist.group.WORLD)
    assert torch.equal(dst["w"].data, full_w.chunk(world, 1)[rank])
    assert torch.equal(dst["b"].data, full_b.chunk(world, 0)[rank])
    return True


def test_reshard_column_to_row_parallel():
    assert all(run_distributed(_reshard, 2))


def test_reshard_plan_prefers_local_replica_and_detects_holes():
    import pytest

    from megatron_b200.core.resharding import ShardDesc, build_reshard_plan

    src = [ShardDesc("w", (4, 4), (0, 0), (4, 4), 0), ShardDesc("w", (4, 4), (0, 0), (4, 4), 1)]
    dst = [ShardDesc("w", (4, 4), (0, 0), (2, 4), 0), ShardDesc("w", (4, 4), (2, 0), (2, 4), 1)]
    plan = build_reshard_plan(src, dst)
    assert all(op.src_rank == op.dst_rank for op in plan), "replicated sources: every rank should copy from itself"
    with pytest.raises(ValueError):
        build_reshard_plan([ShardDesc("w", (4, 4), (0, 0), (2, 4), 0)], [ShardDesc("w", (4, 4), (0, 0), (4, 4), 1)])


def test_cuda_graph_manager_falls_back_to_eager_on_cpu():
    from megatron_b200.core.transformer.cuda_graphs import graph_module

    net = graph_module(torch.nn.Linear(4, 4), warmup_steps=0)
    x = torch.randn(2, 4, requires_grad=True)
    net(x).sum().backward()
    assert x.grad is not None and net.cudagraph_manager.captured == {}


def test_nccl_mem_shim_allocates_plain_memory_without_gpu():
    from megatron_b200.core import nccl_allocator

    with nccl_allocator.nccl_mem(group=None) as mem:
        t = mem.alloc(16, torch.float32)
    assert t.shape == (16,) and float(t.abs().sum()) == 0.0
