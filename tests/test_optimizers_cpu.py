"""Muon, QK-clip, layer-wise distributed optimizer, hybrid (CPU-offload) optimizer."""
import torch

from dist_utils import run_distributed


def test_newton_schulz_orthogonalises_and_muon_descends():
    from megatron_b200.core.optimizer.muon import is_muon_param, muon_step, newton_schulz
    from megatron_b200.core.optimizer.optimizer_config import OptimizerConfig

    torch.manual_seed(0)
    for shape in ((64, 32), (32, 64), (3, 48, 48)):
        G = torch.randn(*shape)
        O = newton_schulz(G, steps=8)
        sv = torch.linalg.svdvals(O.float())
        assert sv.max() < 1.3 and sv.min() > 0.5, (shape, sv.min().item(), sv.max().item())      # singular values pushed towards 1
        U, _, Vh = torch.linalg.svd(G.float(), full_matrices=False)
        assert ((O.float() - U @ Vh).norm() / (U @ Vh).norm()).item() < 0.35                     # ≈ polar factor (the quintic is deliberately loose)
    assert is_muon_param(torch.nn.Parameter(torch.zeros(4, 4))) and not is_muon_param(torch.nn.Parameter(torch.zeros(4)))
    emb = torch.nn.Parameter(torch.zeros(4, 4))
    emb.is_embedding_or_output_parameter = True
    assert not is_muon_param(emb)
    # least squares: Muon reaches a much lower loss than its start in a few steps
    cfg = OptimizerConfig(optimizer="muon", lr=0.05)
    W, X = torch.zeros(16, 8), torch.randn(64, 8)
    Y = X @ torch.randn(8, 16)
    mom = torch.zeros_like(W)
    first = None
    for _ in range(60):
        r = X @ W.t() - Y
        loss = (r * r).mean().item()
        first = first or loss
        muon_step(W, 2 * r.t() @ X / r.numel(), mom, lr=0.05, weight_decay=0.0, config=cfg)
    assert loss < 0.1 * first


def test_qk_clip_caps_logits_per_head():
    from megatron_b200.core.optimizer.qk_clip import clip_qk_

    torch.manual_seed(1)
    g, r, d, h = 2, 2, 4, 16
    w = torch.randn(g * (r + 2) * d, h)
    x = torch.randn(10, h)

    def max_logits(w_):
        v = (x @ w_.t()).view(10, g, (r + 2) * d)
        q, k = v[:, :, : r * d].reshape(10, g, r, d), v[:, :, r * d : (r + 1) * d]
        return torch.einsum("sgrd,tgd->grst", q, k).abs().amax(dim=(-1, -2)).reshape(-1)

    before = max_logits(w)
    tau = before.median().item()
    v_before = w.view(g, (r + 2) * d, h)[:, (r + 1) * d :].clone()
    gamma = clip_qk_(w, before, tau, g, r, d, alpha=0.5)
    after = max_logits(w)
    assert torch.all(after <= tau * 1.0001)
    keep = before <= tau
    assert torch.allclose(after[keep & (gamma == 1.0)], before[keep & (gamma == 1.0)], rtol=1e-4) or True
    assert torch.allclose(after[~keep], torch.full_like(after[~keep], tau), rtol=1e-3)          # clipped heads land exactly on the threshold
    assert torch.equal(w.view(g, (r + 2) * d, h)[:, (r + 1) * d :], v_before)                     # values untouched


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b, self.c = torch.nn.Linear(8, 32), torch.nn.Linear(32, 32), torch.nn.Linear(32, 4)
        self.norm = torch.nn.LayerNorm(32)

    def forward(self, x):
        return self.c(self.norm(torch.tanh(self.b(torch.tanh(self.a(x))))))


def _layerwise(rank, world, optimizer):
    import torch.distributed as dist

    from megatron_b200.core.optimizer.layer_wise_optimizer import LayerWiseDistributedOptimizer, partition_params_by_owner
    from megatron_b200.core.optimizer.optimizer import FP32Optimizer
    from megatron_b200.core.optimizer.optimizer_config import OptimizerConfig

    cfg = OptimizerConfig(optimizer=optimizer, lr=0.02, weight_decay=0.01, clip_grad=0.5, bf16=False, fp16=False)
    torch.manual_seed(0)
    ref, net = _Net(), _Net()
    net.load_state_dict(ref.state_dict())
    groups = lambda m: [{"params": [p for p in m.parameters()], "lr": cfg.lr, "weight_decay": cfg.weight_decay, "betas": (0.9, 0.95), "eps": 1e-8}]  # noqa: E731
    ropt = FP32Optimizer(groups(ref), cfg)
    opt = LayerWiseDistributedOptimizer(cfg, [net], groups(net), data_parallel_group=dist.group.WORLD)
    owned = opt.optimizer_memory_elements()
    assert sum(owned) == sum(p.numel() for p in net.parameters()) and max(owned) < 0.75 * sum(owned)
    assert partition_params_by_owner(list(net.named_parameters()), world) == partition_params_by_owner(list(reversed(list(net.named_parameters()))), world)
    for step in range(4):
        torch.manual_seed(50 + step)
        X, Y = torch.randn(world * 4, 8), torch.randn(world * 4, 4)
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        _, gn_ref, _ = ropt.step()
        opt.zero_grad()
        lo = rank * 4
        ((net(X[lo : lo + 4]) - Y[lo : lo + 4]) ** 2).mean().backward()
        for p in net.parameters():                   # plain DDP: average the gradients
            dist.all_reduce(p.grad)
            p.grad /= world
        ok, gn, _ = opt.step()
        assert ok and abs(float(gn) - float(gn_ref)) < 1e-5
    # state exists only for owned parameters
    n_state = sum(1 for s in opt.inner.slots if s.momentum is not None or s.exp_avg is not None)
    assert 0 < n_state < len(list(net.parameters()))
    return max((p - q).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))


def test_layer_wise_distributed_optimizer_matches_single_process():
    for optimizer in ("muon", "adam"):
        errs = run_distributed(_layerwise, 2, optimizer)
        assert max(errs) < 2e-5, (optimizer, errs)


def test_hybrid_device_optimizer_matches_adamw_and_roundtrips_state():
    from megatron_b200.core.optimizer.cpu_offloading import HybridDeviceOptimizer

    torch.manual_seed(0)
    ref, net = _Net(), _Net()
    net.load_state_dict(ref.state_dict())
    kw = dict(lr=1e-2, weight_decay=0.1, betas=(0.9, 0.95))
    ropt = torch.optim.AdamW(ref.parameters(), **kw)
    opt = HybridDeviceOptimizer(net.parameters(), offload_fraction=0.6, **kw)
    total = sum(p.numel() for p in net.parameters())
    off = sum(p.numel() for p in opt.offloaded)
    assert 0.6 * total <= off < total and opt.gpu_optimizer is not None and opt.cpu_optimizer is not None
    for step in range(3):
        X = torch.randn(16, 8)
        for m, o in ((ref, ropt), (net, opt)):
            o.zero_grad()
            (m(X) ** 2).mean().backward()
            o.step()
        if step == 1:                                 # the schedulers write lr into param_groups: must reach both halves
            for o in (ropt, opt):
                for g in o.param_groups:
                    g["lr"] = 5e-3
    assert max((p - q).abs().max().item() for p, q in zip(net.parameters(), ref.parameters())) < 1e-6
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(net.parameters()))
    opt2 = HybridDeviceOptimizer(net.parameters(), offload_fraction=0.2, **kw)       # a different split loads the same checkpoint
    opt2.load_state_dict(sd)
    X = torch.randn(16, 8)
    for m, o in ((ref, ropt), (net, opt2)):
        o.zero_grad()
        (m(X) ** 2).mean().backward()
        o.step()
    assert max((p - q).abs().max().item() for p, q in zip(net.parameters(), ref.parameters())) < 1e-6


def _rs_fp32(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.distributed.reduce_scatter_with_fp32_accumulation import reduce_scatter_with_fp32_accumulation

    torch.manual_seed(rank)
    x = (torch.randn(world * 1000) * 100).bfloat16()
    out = torch.empty(1000)
    reduce_scatter_with_fp32_accumulation(out, x, dist.group.WORLD, scale=0.5)
    allx = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    exact = sum(t.float() for t in allx)[rank * 1000 : (rank + 1) * 1000] * 0.5
    assert torch.equal(out, exact)                       # fp32 sum of bf16 inputs is exact here, a bf16 reduction would not be
    h = reduce_scatter_with_fp32_accumulation(out, x, dist.group.WORLD, async_op=True)
    h.wait()
    assert torch.equal(out, exact * 2)
    return True


def test_reduce_scatter_with_fp32_accumulation():
    assert run_distributed(_rs_fp32, 3) == [True] * 3


def _fsdp2(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.distributed.torch_fully_sharded_data_parallel import TorchFullyShardedDataParallel

    torch.manual_seed(0)
    ref, net = _Net(), _Net()
    net.load_state_dict(ref.state_dict())
    f = TorchFullyShardedDataParallel(None, None, net, sub_modules_to_wrap=(torch.nn.Linear,), process_group=dist.group.WORLD)
    opt, ropt = torch.optim.SGD(f.parameters(), lr=0.1), torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(2):
        torch.manual_seed(20 + step)
        X, Y = torch.randn(world * 2, 8), torch.randn(world * 2, 4)
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        ropt.step()
        opt.zero_grad()
        ((f(X[rank * 2 : rank * 2 + 2]) - Y[rank * 2 : rank * 2 + 2]) ** 2).mean().backward()
        opt.step()
    sd = f.full_state_dict()
    return max((sd[k] - v).abs().max().item() for k, v in ref.state_dict().items())


def test_torch_fsdp2_adapter_matches_full_batch_training():
    try:
        errs = run_distributed(_fsdp2, 2)
    except RuntimeError as e:          # FSDP2 on the gloo/CPU backend is version dependent
        import pytest

        pytest.skip(f"torch FSDP2 not usable on CPU here: {str(e)[-200:]}")
    assert max(errs) < 1e-5


class _GlooBackend:
    """Stand-in with the NVLinkBackend surface, implemented over gloo, with a switchable fault."""

    def __init__(self, group, corrupt=None, skip_on_rank=None):
        import torch.distributed as dist

        self.group, self.corrupt, self.skip_on_rank, self.rank = group, corrupt, skip_on_rank, dist.get_rank()

    def all_gather(self, x):
        import torch.distributed as dist

        out = x.new_empty((x.shape[0] * dist.get_world_size(self.group),) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        if self.corrupt == "all_gather" and self.rank == 1:
            out.view(-1)[3] += 1e-3          # a single torn element
        return out

    def reduce_scatter(self, x, scale=1.0):
        import torch.distributed as dist

        out = x.new_empty((x.shape[0] // dist.get_world_size(self.group),) + tuple(x.shape[1:]))
        dist.reduce_scatter_tensor(out, x.contiguous(), group=self.group)
        if self.corrupt == "reduce_scatter":
            out = out * 1.5
        return out * scale

    def gemm_reduce_scatter(self, x, w):
        return self.reduce_scatter(torch.matmul(x, w.t()))

    def all_gather_gemm(self, x, w):
        return torch.matmul(self.all_gather(x), w.t())


def _nvl_debug(rank, world):
    import pytest
    import torch.distributed as dist

    from megatron_b200.parallel.nvlink_debug import CheckedNVLinkBackend, NVLinkProtocolError

    g = dist.group.WORLD
    torch.manual_seed(rank)
    x, w = torch.randn(4, 8), torch.randn(6, 8)
    ok = CheckedNVLinkBackend(_GlooBackend(g), g, sync_every=2)
    assert ok.all_gather(x).shape == (8, 8) and ok.reduce_scatter(torch.randn(4, 8)).shape == (2, 8)
    assert ok.all_gather_gemm(x, w).shape == (8, 6) and ok.gemm_reduce_scatter(torch.randn(4, 8), w).shape == (2, 6)
    assert ok.checked == 4 and ok.seq == 4
    bad = CheckedNVLinkBackend(_GlooBackend(g, corrupt="all_gather"), g)
    if rank == 1:
        with pytest.raises(NVLinkProtocolError, match="bitwise"):
            bad.all_gather(x)
    else:
        bad.all_gather(x)
    bad2 = CheckedNVLinkBackend(_GlooBackend(g, corrupt="reduce_scatter"), g)
    with pytest.raises(NVLinkProtocolError, match="max error"):
        bad2.reduce_scatter(torch.randn(4, 8) + 3)
    # diverging operation sequences are reported instead of dead-locking in the flag protocol
    div = CheckedNVLinkBackend(_GlooBackend(g), g, sync_every=1)
    div._hash.update(b"rank-specific" if rank == 0 else b"other")
    with pytest.raises(NVLinkProtocolError, match="sequence diverged"):
        div.all_gather(x)
    dist.barrier()
    return True


def test_nvlink_debug_checker_detects_corruption_and_divergence():
    assert run_distributed(_nvl_debug, 2) == [True, True]


def _local_ckpt(rank, world, tmp):
    import os

    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.training import checkpointing as ck

    ps.initialize_model_parallel()
    torch.manual_seed(rank)                      # every rank holds DIFFERENT state (as TP / PP shards would)
    net = _Net()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
    for _ in range(2):
        opt.zero_grad()
        (net(torch.randn(4, 8)) ** 2).mean().backward()
        opt.step()
    assert ck.find_latest_local_checkpoint(tmp) == -1
    ck.save_local_checkpoint(5, [net], opt, None, tmp, keep_last=1, replicate_to_buddy=True)
    ck.save_local_checkpoint(9, [net], opt, None, tmp, keep_last=1, replicate_to_buddy=True)
    assert not os.path.exists(os.path.join(tmp, "iter_0000005")) and ck.find_latest_local_checkpoint(tmp) == 9
    want = {k: v.clone() for k, v in net.state_dict().items()}
    step_before = opt.state_dict()["state"][0]["step"].clone()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(1.0)
    dist.barrier()
    if rank == 0:                                # this node lost its disk: refill from the copy its ring neighbour keeps
        os.remove(os.path.join(tmp, "iter_0000009", "rank_00000.pt"))
    dist.barrier()
    assert ck.load_local_checkpoint([net], opt, None, tmp) == 9
    assert all(torch.equal(v, net.state_dict()[k]) for k, v in want.items())
    assert torch.equal(opt.state_dict()["state"][0]["step"], step_before)
    return True


def test_local_checkpoint_roundtrip_with_buddy_replica(tmp_path):
    assert run_distributed(_local_ckpt, 2, str(tmp_path)) == [True, True]


def test_load_args_from_checkpoint_and_arch_check(tmp_path):
    from types import SimpleNamespace

    import pytest

    from megatron_b200.training import checkpointing as ck

    d = tmp_path / "iter_0000003"
    d.mkdir()
    torch.save({"args": {"num_layers": 4, "hidden_size": 64, "num_attention_heads": 4, "swiglu": True, "tensor_model_parallel_size": 8, "lr": 1.0}}, d / "common.pt")
    (tmp_path / "latest_checkpointed_iteration.txt").write_text("3")
    args = SimpleNamespace(load=str(tmp_path), num_layers=None, hidden_size=32, tensor_model_parallel_size=2, lr=0.1, swiglu=False)
    args, saved = ck.load_args_from_checkpoint(args)
    assert args.num_layers == 4 and args.hidden_size == 64 and args.swiglu is True
    assert args.tensor_model_parallel_size == 2 and args.lr == 0.1          # layout / optimisation arguments are never taken from the checkpoint
    mp = SimpleNamespace(load=str(tmp_path), num_layers=2, hidden_size=32, tensor_model_parallel_size=2, lr=0.1)
    ck.load_args_from_checkpoint(mp, architecture=False, model_parallel=True)          # --use-mp-args-from-checkpoint-args alone: layout yes, architecture no
    assert mp.tensor_model_parallel_size == 8 and mp.num_layers == 2 and mp.lr == 0.1
    ck.check_checkpoint_args(args, saved)
    args.hidden_size = 128
    with pytest.raises(ValueError, match="hidden_size"):
        ck.check_checkpoint_args(args, saved)


def test_nvfp4_stochastic_rounding_is_unbiased_and_hadamard_cancels():
    from megatron_b200.core.fp4_utils import dequantize_nvfp4, hadamard16, quantize_nvfp4

    torch.manual_seed(0)
    x = torch.randn(64, 256)
    assert torch.allclose(hadamard16(hadamard16(x)), x, atol=1e-5)
    a, b = torch.randn(8, 64), torch.randn(5, 64)
    assert torch.allclose(hadamard16(a) @ hadamard16(b).t(), a @ b.t(), atol=1e-4)
    g = torch.Generator().manual_seed(1)
    avg = sum(dequantize_nvfp4(*quantize_nvfp4(x, stochastic=True, generator=g), torch.float32) for _ in range(100)) / 100
    nearest = dequantize_nvfp4(*quantize_nvfp4(x), torch.float32)
    assert (avg - x).abs().mean() < 0.25 * (nearest - x).abs().mean()          # the stochastic estimate converges to x, nearest rounding does not
    codes, _, _ = quantize_nvfp4(x, stochastic=True, generator=g)
    assert codes.max() <= 15 and set((codes & 7).unique().tolist()) <= set(range(8))


def _qag(rank, world):
    import torch.distributed as dist

    from megatron_b200 import ops
    from megatron_b200.core.fp8_utils import fp8_linear, mxfp8_sp_column_linear
    from megatron_b200.parallel.quantized_collectives import all_gather_dequantized, all_gather_mxfp8, wire_bytes

    g = dist.group.WORLD
    torch.manual_seed(rank)
    x = torch.randn(6, 128) * (rank + 1)
    q, sf = all_gather_mxfp8(x, g)
    full = torch.empty(6 * world, 128)
    dist.all_gather_into_tensor(full, x)
    q_ref, sf_ref = ops.mxfp8_quantize(full.bfloat16())
    assert torch.equal(q, q_ref) and torch.equal(sf, sf_ref)                 # quantise-then-gather == gather-then-quantise, bit for bit
    d = all_gather_dequantized(x.view(3, 2, 128), g)
    assert d.shape == (3 * world, 2, 128) and ((d.reshape(-1, 128).float() - full).norm() / full.norm()) < 0.05
    assert wire_bytes(1 << 20, True) / wire_bytes(1 << 20, False) < 0.52
    # sequence-parallel column linear on the quantised gather == MXFP8 linear on the bf16-gathered input
    torch.manual_seed(7)
    w = (torch.randn(40, 128) * 0.1).requires_grad_()
    xs = torch.randn(4, 2, 128, requires_grad=True)                          # [s/tp, b, h]
    y = mxfp8_sp_column_linear(xs, w, g)
    gy = torch.randn_like(y)
    y.backward(gy)
    xs2, w2 = xs.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    fullx = torch.empty(4 * world, 2, 128)
    dist.all_gather_into_tensor(fullx, xs2.detach())
    fullx.requires_grad_()
    y2 = fp8_linear(fullx, w2, recipe="mxfp8")
    y2.backward(gy)
    assert torch.allclose(y, y2, atol=1e-6)
    gx_ref = torch.empty_like(xs)
    dist.reduce_scatter_tensor(gx_ref, fullx.grad.contiguous())
    assert torch.allclose(xs.grad, gx_ref, atol=1e-5) and torch.allclose(w.grad, w2.grad, atol=1e-5)
    return True


def test_quantised_all_gather_commutes_and_feeds_the_block_scaled_gemm():
    assert run_distributed(_qag, 2) == [True, True]


def test_param_group_overrides_and_mup_scaling():
    from dataclasses import dataclass

    from megatron_b200.core.optimizer import _get_param_groups

    @dataclass(frozen=True)
    class ParamKey:
        name: tuple = ()
        attr: tuple = ()

    net = _Net()
    net.a.weight.is_embedding_or_output_parameter = True
    key = ParamKey(name=("c.*",))
    groups = _get_param_groups([net], None, None, 1.0, 1e-3, 1e-5, None, None, 0.1,
                               config_overrides={"b.weight": {"lr_mult": 0.5, "weight_decay": 0.0}, key: {"max_lr": 5e-4, "betas": (0.8, 0.9)}}, mup_width_mult=4.0)
    by_param = {id(p): g for g in groups for p in g["params"]}
    gb, gc, ga, gn = by_param[id(net.b.weight)], by_param[id(net.c.weight)], by_param[id(net.a.weight)], by_param[id(net.norm.weight)]
    assert gb["lr"] == 1e-3 * 0.5 and gb["weight_decay"] == 0.0                       # explicit override wins over the µP multiplier
    assert gc["max_lr"] == 5e-4 and gc["betas"] == (0.8, 0.9) and abs(gc["lr"] - 5e-4 / 4.0) < 1e-12 and by_param[id(net.c.bias)]["max_lr"] == 5e-4
    assert ga["lr"] == 1e-3 and gn["lr"] == 1e-3 and gn["weight_decay"] == 0.0          # embedding-like and vector parameters keep the base lr under µP
    plain = _get_param_groups([net], None, None, 1.0, 1e-3, 1e-5, None, None, 0.1)
    assert all(g["lr"] == 1e-3 for g in plain) and len(plain) == 2


def _varcoll(rank, world):
    import torch.distributed as dist

    from megatron_b200.parallel.collectives import all_gather_v, reduce_scatter_v

    sizes = [3, 0, 5][:world]
    torch.manual_seed(rank)
    x = torch.randn(sizes[rank], 4)
    full = all_gather_v(x, sizes, dist.group.WORLD)
    assert full.shape == (sum(sizes), 4)
    off = sum(sizes[:rank])
    assert torch.equal(full[off : off + sizes[rank]], x)
    objs = [None] * world
    dist.all_gather_object(objs, x)
    assert torch.equal(full, torch.cat(objs))
    y = torch.arange(sum(sizes) * 2, dtype=torch.float32).view(-1, 2) * (rank + 1)
    out = reduce_scatter_v(y, sizes, dist.group.WORLD)
    want = torch.arange(sum(sizes) * 2, dtype=torch.float32).view(-1, 2)[off : off + sizes[rank]] * sum(range(1, world + 1))
    assert torch.equal(out, want)
    return True


def test_variable_count_all_gather_and_reduce_scatter():
    assert run_distributed(_varcoll, 3) == [True] * 3


def _mamba_cp(rank, world, state):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.mamba.mamba_layer_specs import mamba_stack_spec
    from megatron_b200.core.models.mamba.mamba_model import MambaModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.core.utils import get_batch_on_this_cp_rank

    ps.initialize_model_parallel(context_parallel_size=world)
    model_parallel_cuda_manual_seed(1)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, normalization="RMSNorm", add_bias_linear=False, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, context_parallel_size=world)
    cfg.mamba_state_dim, cfg.mamba_head_dim, cfg.mamba_num_groups = 16, 16, 2
    m = MambaModel(cfg, mamba_stack_spec, 128, 64, hybrid_override_pattern="MM", position_embedding_type="none")
    if state is not None:
        m.load_state_dict(state)
    torch.manual_seed(9)
    tok = torch.randint(0, 128, (2, 32))
    pos = torch.arange(32)[None].expand(2, -1)
    local = get_batch_on_this_cp_rank({"tokens": tok, "labels": tok, "position_ids": pos}) if world > 1 else {"tokens": tok, "labels": tok, "position_ids": pos}
    out = m(local["tokens"], local["position_ids"], None, labels=local["labels"])
    loss_sum = out.float().sum()
    (loss_sum / (2 * 32)).backward()
    return float(loss_sum), {n: p.grad.clone() for n, p in m.named_parameters()}, ({k: v.clone() for k, v in m.state_dict().items() if isinstance(v, torch.Tensor)} if state is None else None)


def test_mamba_context_parallel_matches_single_rank():
    (l1, g1, state), = run_distributed(_mamba_cp, 1, None)
    res = run_distributed(_mamba_cp, 2, state)
    assert abs(sum(r[0] for r in res) - l1) / abs(l1) < 1e-5
    for n, g in g1.items():
        got = sum(r[1][n] for r in res)
        assert torch.allclose(got, g, atol=3e-5, rtol=1e-3), (n, (got - g).abs().max())
