"""Megatron-FSDP layer (core/distributed/fsdp/src/megatron_fsdp): buffer index math, bucket allocators, the two pipelines,
and end-to-end equivalence of every sharding strategy with plain single-process training."""
import copy

import pytest
import torch

from dist_utils import run_distributed


def test_buffer_index_alignment_and_uneven_item_slices():
    from megatron_b200.core.distributed.fsdp.src.megatron_fsdp.param_and_grad_buffer import ALIGN_BYTES, build_data_parallel_buffer_index

    shapes = [torch.Size([5, 7]), torch.Size([3]), torch.Size([11, 2])]
    for world in (1, 2, 3, 8):
        per_rank = []
        for r in range(world):
            items, bucket, shard = build_data_parallel_buffer_index(shapes, r, world, True, dtype=torch.bfloat16)
            assert bucket.size % (world * (ALIGN_BYTES // 2)) == 0 and shard.size * world == bucket.size
            assert all(it.global_data_index % 8 == 0 for it in items), "items start on 16-byte boundaries"
            per_rank.append((shard.global_data_index, shard.size))
        assert [p[0] for p in per_rank] == [r * per_rank[0][1] for r in range(world)]


def test_allocators_recycle_and_keep_addresses():
    from megatron_b200.core.distributed.fsdp.src.megatron_fsdp.param_and_grad_buffer import (
        FixedPoolAllocator, MaxPoolAllocator, ParameterGroup, RotaryBucketAllocator, StorageResizeBasedBucketAllocator)

    P = lambda *shape: torch.nn.Parameter(torch.zeros(*shape))   # noqa: E731
    # four identical units of two buckets each + one odd group outside any unit
    groups = [ParameterGroup([P(4, 4)], torch.float32, fsdp_unit_id=None)]
    for u in range(4):
        groups += [ParameterGroup([P(8, 8)], torch.float32, fsdp_unit_id=u), ParameterGroup([P(8)], torch.float32, fsdp_unit_id=u)]
    fp = FixedPoolAllocator("w", groups, size=2)
    a1 = fp.allocate(1, 64, torch.float32, "cpu").data.data_ptr()
    a3 = fp.allocate(3, 64, torch.float32, "cpu").data.data_ptr()
    assert a1 != a3 and not fp.can_allocate(5), "unit 2 shares group 0 with unit 0, which is busy"
    fp.free(1)
    assert fp.can_allocate(5) and fp.allocate(5, 64, torch.float32, "cpu").data.data_ptr() == a1, "unit 2 reuses unit 0's memory"
    assert not fp.can_allocate(7), "unit 3 must land where unit 1 lives (saved autograd views alias it), never in a free group"
    assert fp.allocate(7, 64, torch.float32, "cpu").data.numel() == 64 and fp.pool_misses == 1, "non-strict (gradient) pools take the slow path"
    with pytest.raises(RuntimeError):
        strict = FixedPoolAllocator("w", groups, size=2, strict=True)
        strict.allocate(1, 64, torch.float32, "cpu")
        strict.allocate(5, 64, torch.float32, "cpu")
    odd = fp.allocate(0, 16, torch.float32, "cpu")
    fp.free(0)
    assert odd.data.untyped_storage().size() == 0, "non-unit buckets fall back to storage resizing"
    # heterogeneous units: bucket sizes differ per unit, the pool holds per-rank maxima
    het = [ParameterGroup([P(10)], torch.float32, fsdp_unit_id=0), ParameterGroup([P(100)], torch.float32, fsdp_unit_id=0),
           ParameterGroup([P(50)], torch.float32, fsdp_unit_id=1), ParameterGroup([P(60)], torch.float32, fsdp_unit_id=1)]
    mp = MaxPoolAllocator("w", het, size=2)
    assert mp.max_dtype_bucket_sizes[torch.float32] == [50, 100]
    assert mp.allocate(0, 10, torch.float32, "cpu").data.numel() == 10 and mp.allocate(3, 60, torch.float32, "cpu").data.numel() == 60
    ro = RotaryBucketAllocator("g")
    x = ro.allocate(0, 32, torch.float32, "cpu").data.data_ptr()
    ro.allocate(1, 16, torch.float32, "cpu")
    ro.free(0)
    assert ro.allocate(2, 8, torch.bfloat16, "cpu").data.data_ptr() == x, "the lowest idle slot is reused, whatever the dtype"
    sr = StorageResizeBasedBucketAllocator()
    t = sr.allocate(0, 8, torch.float32, "cpu").data
    sr.free(0)
    assert t.untyped_storage().size() == 0
    assert sr.allocate(0, 8, torch.float32, "cpu").data is t and t.untyped_storage().size() == 32


class _Block(torch.nn.Module):
    def __init__(self, d, dtype):
        super().__init__()
        self.norm = torch.nn.LayerNorm(d, dtype=dtype)
        self.fc = torch.nn.Linear(d, d, dtype=dtype)

    def forward(self, x):
        return x + torch.tanh(self.fc(self.norm(x)))


class _Net(torch.nn.Module):
    def __init__(self, d=12, n=3, dtype=torch.float32):
        super().__init__()
        self.inp = torch.nn.Linear(5, d, dtype=dtype)
        self.blocks = torch.nn.ModuleList([_Block(d, dtype) for _ in range(n)])
        self.unused = torch.nn.Parameter(torch.ones(3, dtype=dtype))       # never receives a gradient
        self.head = torch.nn.Linear(d, 2, dtype=dtype)

    def forward(self, x):
        x = self.inp(x)
        for b in self.blocks:
            x = b(x)
        return self.head(x)


def _train_reference(model, batches, lr, steps, compute_dtype=torch.float32):
    """fp32 master weights + (optionally) a low-precision compute copy — what the FSDP wrapper does, in one process."""
    opt = (torch.optim.Adam if compute_dtype == torch.float32 else torch.optim.SGD)(model.parameters(), lr=lr)
    low = copy.deepcopy(model).to(compute_dtype) if compute_dtype != torch.float32 else None
    for s in range(steps):
        opt.zero_grad()
        if low is not None:
            low.load_state_dict({k: v.to(compute_dtype) for k, v in model.state_dict().items()})
            low.zero_grad()
        for mb in batches[s]:                                  # micro-batches: gradients of the MEAN over all samples of the step
            m = low if low is not None else model
            (m(mb.to(compute_dtype)).float().pow(2).mean() / len(batches[s])).backward()
        if low is not None:
            for p, q in zip(model.parameters(), low.parameters()):
                p.grad = q.grad.float() if q.grad is not None else None
        opt.step()
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _fsdp_worker(rank, world, strategy, dtype_name, allocator, n_micro):
    import torch.distributed as dist

    from megatron_b200.core.distributed.distributed_data_parallel_config import DistributedDataParallelConfig
    from megatron_b200.core.distributed.fsdp.src.megatron_fsdp import fully_shard
    from megatron_b200.core.distributed.fsdp.src.megatron_fsdp.param_and_grad_buffer import FixedPoolAllocator

    dtype = getattr(torch, dtype_name)
    torch.manual_seed(0)
    base = _Net(dtype=dtype)
    steps, lr = 3, (1e-2 if dtype == torch.float32 else 0.1)
    torch.manual_seed(1)
    # step -> micro-batch -> rank -> samples ; the reference sees the concatenation over ranks
    data = [[[torch.randn(4, 5, dtype=dtype) for _ in range(world)] for _ in range(n_micro)] for _ in range(steps)]
    ref_model = copy.deepcopy(base).float()
    want = _train_reference(ref_model, [[torch.cat(mb).float() for mb in step] for step in data], lr, steps, dtype)

    ddp_cfg = DistributedDataParallelConfig(data_parallel_sharding_strategy=strategy, fsdp_double_buffer=(allocator == "fixed"))
    model = copy.deepcopy(base)
    opt = (torch.optim.Adam if dtype == torch.float32 else torch.optim.SGD)(model.parameters(), lr=lr)
    model, opt = fully_shard(model, opt, fsdp_unit_modules=[_Block], zero_dp_strategy=strategy, dp_shard_group=dist.group.WORLD, ddp_config=ddp_cfg,
                             fsdp_double_buffer=(allocator == "fixed"))
    buf = model.param_and_grad_buffer
    if strategy == "optim_grads_params":
        assert all(p.numel() == 0 for p in model.module.parameters()), "ZeRO-3: no full parameter is resident between steps"
        resident = sum(g.model_weight_buffer.data.numel() for g in buf.parameter_groups)
        total = sum(g.model_weight_buffer.bucket_index.size for g in buf.parameter_groups)
        assert resident * world == total
    for s in range(steps):
        opt.zero_grad()
        for i, mb in enumerate(data[s]):
            ctx = model.no_sync() if i + 1 < n_micro else torch.enable_grad()
            with ctx:
                (model(mb[rank]).float().pow(2).mean() / n_micro).backward()
        opt.step()
    got = model.gather_full_state_dict()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for k, v in want.items():
        assert torch.allclose(got[k].float(), v, atol=tol, rtol=tol), (strategy, k, float((got[k].float() - v).abs().max()))
    assert torch.equal(got["unused"].float(), torch.ones(3)), "a parameter without gradient is left alone"
    info = {"launches": list(model.all_gather_pipeline.launch_log), "reduced": list(model.grad_reduce_pipeline.reduced_log)}
    if allocator == "fixed" and strategy == "optim_grads_params":
        assert isinstance(buf.weight_alloc, FixedPoolAllocator)
        info["pool_entries"] = len(buf.weight_alloc.pool)
        info["misses"] = buf.weight_alloc.pool_misses
    return info


@pytest.mark.parametrize("strategy", ["no_shard", "optim", "optim_grads", "optim_grads_params"])
def test_every_strategy_matches_single_process_adam(strategy):
    run_distributed(_fsdp_worker, 2, strategy, "float32", "auto", 2)


def test_zero3_uneven_world_bf16_main_weights_and_double_buffer():
    out = run_distributed(_fsdp_worker, 3, "optim_grads_params", "bfloat16", "fixed", 1)
    info = out[0]
    # 3 steps x (forward + backward) gathers of the three block units, in pass order; two buckets per unit (fc+norm share dtype -> one)
    assert info["pool_entries"] <= 2, "the double buffer never holds more than two unit buckets"
    assert len(info["launches"]) >= 3 * 2 * 3


def _uneven_worker(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.distributed.fsdp.src.megatron_fsdp.uneven_dtensor import (
        gather_uneven_dtensor_to_full_tensor, split_dtensor, update_uneven_dtensor_chunk_metadata, validate_uneven_dtensor)

    full = torch.arange(35.0).view(5, 7)
    sh = split_dtensor(full, [0, 20, 20, 35], rank)          # rank 1 owns nothing
    validate_uneven_dtensor(sh, dist.group.WORLD)
    assert torch.equal(gather_uneven_dtensor_to_full_tensor(sh, dist.group.WORLD), full)
    if rank == 2:
        update_uneven_dtensor_chunk_metadata(sh, 19, 34)      # overlaps rank 0's chunk
    try:
        validate_uneven_dtensor(sh, dist.group.WORLD)
        return False
    except ValueError:
        return True


def test_uneven_shard_helpers():
    assert all(run_distributed(_uneven_worker, 3))
