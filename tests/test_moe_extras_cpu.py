"""MoE side systems: router replay / trace, metrics tracker, dense→MoE upcycling, paged stash, serving dispatchers (CPU, gloo)."""
import os

import pytest
import torch

from dist_utils import run_distributed

_KW = dict(use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, add_bias_linear=False)


def _init(seed=1, tp=1, ep=1):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(tp, 1, expert_model_parallel_size=ep)
    model_parallel_cuda_manual_seed(seed)


def _moe_layer(grouped=False, **kw):
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    base = dict(num_layers=1, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, num_moe_experts=4, moe_ffn_hidden_size=32,
                moe_router_topk=2, moe_grouped_gemm=grouped, gated_linear_unit=True, activation_func=torch.nn.functional.silu, **_KW)
    base.update(kw)
    cfg = TransformerConfig(**base)
    spec = get_gpt_layer_local_spec(num_experts=cfg.num_moe_experts, moe_grouped_gemm=grouped, normalization="RMSNorm")
    from megatron_b200.core.transformer.spec_utils import build_module

    mlp = build_module(spec.submodules.mlp, config=cfg)
    mlp.set_layer_number(1)
    return cfg, mlp


def _replay(rank, world):
    _init()
    from megatron_b200.core.transformer.moe.router_replay import RouterReplay, RouterReplayAction

    RouterReplay.clear_global_router_replay_instances()
    cfg, mlp = _moe_layer(moe_enable_routing_replay=True)
    assert len(RouterReplay.global_router_replay_instances) == 1
    x = torch.randn(6, 2, 32)
    RouterReplay.set_global_router_replay_action(RouterReplayAction.RECORD)
    y0, _ = mlp(x)
    rec = RouterReplay.get_recorded_data()
    assert rec[0].shape == (12, 2)
    # replay a *different* routing: every token to experts (3, 0) — the map follows the replayed ids, probs are their gathered scores
    forced = torch.tensor([[3, 0]]).expand(12, 2).contiguous()
    RouterReplay.set_replay_data([forced])
    RouterReplay.set_global_router_replay_action(RouterReplayAction.REPLAY_FORWARD)
    probs, rmap = mlp.router(x)
    assert rmap[:, [0, 3]].all() and not rmap[:, [1, 2]].any()
    assert torch.allclose(probs.sum(-1), torch.ones(12), atol=1e-5)
    # the recompute inside backward pops the same ids
    RouterReplay.set_global_router_replay_action(RouterReplayAction.REPLAY_BACKWARD)
    probs_b, rmap_b = mlp.router(x)
    assert torch.equal(rmap, rmap_b) and torch.allclose(probs, probs_b)
    with pytest.raises(RuntimeError):
        mlp.router(x)                                           # queue is empty now
    # replaying the recorded ids reproduces the recorded forward exactly
    RouterReplay.set_replay_data(rec)
    RouterReplay.set_global_router_replay_action(RouterReplayAction.REPLAY_FORWARD)
    y1, _ = mlp(x)
    assert torch.allclose(y0, y1, atol=1e-6)
    RouterReplay.clear_global_router_replay_action()
    RouterReplay.clear_global_router_replay_instances()
    return True


def test_router_replay_records_and_forces_routing():
    run_distributed(_replay, 1)


def _trace(rank, world, tmp):
    _init()
    from megatron_b200.core.transformer.moe import router_trace as rt

    cfg, mlp = _moe_layer()
    holder = torch.nn.Module()
    holder.decoder = torch.nn.Module()
    holder.decoder.layers = torch.nn.ModuleList([torch.nn.Module()])
    holder.decoder.layers[0].mlp = mlp
    tr = rt.init_moe_router_tracer(tmp, save_hidden_states=True, save_logits=True, flush_every=2, start_step=1)
    tr.register_hooks(holder)
    x = torch.randn(5, 2, 32)
    mlp(x)                                   # step 0: before start_step → not traced
    tr.advance_step()
    mlp(x)
    mlp(x)                                   # second micro-batch of step 1
    tr.advance_step()
    tr.remove_hooks()
    mlp(x)
    idx = tr.read_index()
    assert [(r["step"], r["microbatch"], r["block"], r["layer"]) for r in idx] == [(1, 0, "decoder", 0), (1, 1, "decoder", 0)]
    top = rt.load_indices_for_record(idx[0], tr.trace_dir)
    probs, rmap = mlp.router(x)
    want = torch.where(rmap, probs, torch.full_like(probs, -1.0)).topk(2, dim=-1).indices
    assert torch.equal(top.long(), want)
    assert rt.load_hidden_states_for_record(idx[0], tr.trace_dir).shape == (10, 32)
    assert rt.load_logits_for_record(idx[1], tr.trace_dir).shape == (10, 4)
    assert rt._parse_router_module_name("mtp.layers.1.transformer_layer.mlp.router") == ("mtp", 1, 0)
    return True


def test_router_tracer_writes_and_reloads_records(tmp_path):
    run_distributed(_trace, 1, str(tmp_path))


def _metrics(rank, world):
    _init()
    from megatron_b200.core.transformer.moe.moe_logging import MoEMetricsTracker, destroy_moe_metrics_tracker, get_moe_metrics_tracker, set_moe_metrics_tracker

    tr = MoEMetricsTracker()
    set_moe_metrics_tracker(tr)
    assert get_moe_metrics_tracker() is tr
    import torch.distributed as dist

    # each rank is one "pipeline stage" holding one of the two MoE layers; values differ per DP replica → averaged
    tr.record("load_balancing_loss", torch.tensor(1.0 + rank), layer_number=rank + 1, num_layers=2, needs_dp_avg=False)
    tr.record("z_loss", torch.tensor(float(rank)), layer_number=1, num_layers=2, avg_group=dist.group.WORLD, needs_dp_avg=False, percentiles=[0.5])

    class W:
        def __init__(self):
            self.s = {}

        def add_scalar(self, k, v, it):
            self.s[k] = v

    w, total = W(), {}
    text = tr.report(loss_scale=0.5, iteration=3, writer=w, total_loss_dict=total, per_layer_logging=True, num_layers=2, pp_group=dist.group.WORLD)
    assert abs(w.s["load_balancing_loss"] - 0.5 * (1.0 + 2.0) / 2) < 1e-6                   # mean over the two layers
    assert abs(w.s["moe/load_balancing_loss_layer_1"] - 1.0) < 1e-6
    assert abs(w.s["z_loss"] - 0.5 * ((0.0 + 1.0) * 2 / 2) / 2) < 1e-6                      # PP-sum of both ranks' layer-0 value, then mean over the group
    assert "load_balancing_loss:" in text and float(total["z_loss"]) > 0
    assert all(float(e.values.abs().sum()) == 0 for e in tr.metrics.values())               # cleared after the report
    destroy_moe_metrics_tracker()
    return True


def test_moe_metrics_tracker_reduces_and_reports():
    run_distributed(_metrics, 2)


def _upcycle(rank, world, grouped):
    _init()
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.moe.upcycling_utils import upcycle_state_dict
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    common = dict(num_layers=2, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, gated_linear_unit=False, activation_func=torch.nn.functional.relu, **_KW)
    dense = GPTModel(TransformerConfig(**common), get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=16)
    # granularity 2 (expert ffn 32), expansion 1 → 2 experts, top-2: the upcycled model computes exactly what the dense one did
    mcfg = TransformerConfig(**common, num_moe_experts=2, moe_ffn_hidden_size=32, moe_router_topk=2, moe_grouped_gemm=grouped)
    moe = GPTModel(mcfg, get_gpt_layer_local_spec(num_experts=2, moe_grouped_gemm=grouped, normalization="RMSNorm"), vocab_size=64, max_sequence_length=16)
    sd = upcycle_state_dict([moe], [dense])["model"]
    moe.load_state_dict(sd)
    tok = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16)[None].expand(2, -1)
    dense.eval(), moe.eval()
    with torch.no_grad():
        a, b = dense(tok, pos, None), moe(tok, pos, None)
    assert (a - b).abs().max().item() < 1e-4, (a - b).abs().max().item()
    # granularity 2, expansion 2 → 4 experts; experts (0,1) and (2,3) hold the two shards and share a router row
    mcfg4 = TransformerConfig(**common, num_moe_experts=4, moe_ffn_hidden_size=32, moe_router_topk=2, moe_grouped_gemm=grouped)
    moe4 = GPTModel(mcfg4, get_gpt_layer_local_spec(num_experts=4, moe_grouped_gemm=grouped, normalization="RMSNorm"), vocab_size=64, max_sequence_length=16)
    sd4 = upcycle_state_dict(moe4, dense)["model"]
    r = sd4["decoder.layers.0.mlp.router.weight"]
    assert torch.equal(r[0], r[1]) and torch.equal(r[2], r[3])
    moe4.load_state_dict(sd4)
    moe4.eval()
    with torch.no_grad():
        c = moe4(tok, pos, None)
    assert (a - c).abs().max().item() < 1e-4               # any complete copy reproduces the dense FFN
    return True


@pytest.mark.parametrize("grouped", [False, True])
def test_upcycled_moe_reproduces_dense_model(grouped):
    run_distributed(_upcycle, 1, grouped)


def test_paged_stash_allocator_and_autograd_roundtrip():
    from megatron_b200.core.transformer.moe import paged_stash as ps_

    buf = ps_.PagedStashBuffer(num_tokens=64, hidden_size=8, page_size=8, device="cpu", dtype=torch.float32)
    assert buf.num_pages == 8
    a, b = torch.randn(24, 8), torch.randn(32, 8)
    pa, pb = ps_.PagedTensor(a.clone(), torch.tensor(19)), ps_.PagedTensor(b.clone(), torch.tensor(32))
    pa.offload_to_stash(buf)
    pb.offload_to_stash(buf)
    assert buf.free_pages() == 8 - 3 - 4 and int(buf.overflow) == 0
    ra = pa.reload_from_stash(buf)
    assert torch.equal(ra[:19], a[:19]) and float(ra[19:].abs().sum()) == 0     # rows past the valid count are not kept
    assert buf.free_pages() == 4
    pc = ps_.PagedTensor(torch.randn(40, 8), torch.tensor(40))                  # 5 pages > 4 free → overflow, nothing allocated
    pc.offload_to_stash(buf)
    assert int(buf.overflow) == 1 and buf.free_pages() == 4
    assert torch.equal(pb.reload_from_stash(buf), b)
    assert buf.free_pages() == 8
    # pages come back in a different order; a second round still round-trips
    buf.overflow.zero_()
    pd = ps_.PagedTensor(b.clone(), torch.tensor(32))
    pd.offload_to_stash(buf)
    assert torch.equal(pd.reload_from_stash(buf), b)

    # through autograd: activations saved inside the context live in the pool between forward and backward
    m = ps_.PagedStashManager.get_instance()
    m.allocate_stash_buffers({(16, torch.float32): 256}, page_size=16, device="cpu")
    w1, w2 = torch.randn(16, 16, requires_grad=True), torch.randn(16, 16, requires_grad=True)
    x = torch.randn(48, 16)

    def f():
        return (torch.relu(x @ w1) @ w2).square().sum()

    f().backward()
    g1, g2 = w1.grad.clone(), w2.grad.clone()
    w1.grad = w2.grad = None
    with ps_.get_paged_stash_context(True):
        loss = f()
    pool = m.buffers[(16, torch.float32)]
    assert m.stashed >= 1 and pool.free_pages() < pool.num_pages
    loss.backward()
    assert torch.allclose(w1.grad, g1) and torch.allclose(w2.grad, g2)
    assert pool.free_pages() == pool.num_pages and not ps_.check_paged_stash_overflow()

    # runner: a step that overflows is re-run without the pool and yields the right gradients
    m.allocate_stash_buffers({(16, torch.float32): 16}, page_size=16, device="cpu")          # far too small
    w1.grad = w2.grad = None
    calls = []

    def fb(data_iterator=None):
        calls.append(next(data_iterator))
        with ps_.get_paged_stash_context(True):
            loss = f()
        loss.backward()
        return loss

    def zero():
        w1.grad = w2.grad = None

    runner = ps_.PagedStashRunner(fb, zero)
    runner(data_iterator=iter([7, 8]))
    assert runner.reruns == 1 and calls == [7, 7]
    assert torch.allclose(w1.grad, g1) and torch.allclose(w2.grad, g2)
    ps_.paged_stash_reset(enabled=False)


def _serving_dispatch(rank, world, kind, grouped):
    _init(ep=world)
    from megatron_b200.core.transformer.moe.token_dispatcher_inference import InferenceAllGatherDispatcherBase, NVLSAllGatherVDispatcher

    torch.manual_seed(5)
    cfg, mlp = _moe_layer(grouped=grouped, moe_token_dispatcher_type="allgather")
    # same expert weights as a single-process reference: rebuild the full set from the seed on every rank
    mlp.eval()
    T = 6 if kind == "nccl" else (6 if rank == 0 else 3)          # nvls: ranks hold different token counts
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(T, 1, 32, generator=g)
    with torch.no_grad():
        # reference: the training dispatcher needs equal counts → pad with rows that are sliced off again (tokens are independent)
        xp = torch.cat([x, torch.zeros(6 - T, 1, 32)]) if T < 6 else x
        want = mlp(xp)[0][:T]
        if kind == "nvls":
            NVLSAllGatherVDispatcher.allocate_buffers(world, max_tokens_per_rank=8)
            NVLSAllGatherVDispatcher.set_real_token_count_tensor(torch.tensor(T, dtype=torch.int32))
        else:
            InferenceAllGatherDispatcherBase.allocate_valid_tokens_tensor()
        mlp.set_inference_dispatcher(kind)
        got, _ = mlp(x)
    assert (want - got).abs().max().item() < 1e-5, (want - got).abs().max().item()
    total = int(InferenceAllGatherDispatcherBase._valid_tokens())
    assert total == (6 * world if kind == "nccl" else 9)
    mlp.set_inference_dispatcher(None)
    NVLSAllGatherVDispatcher._delete_buffers()
    return True


@pytest.mark.parametrize("kind,grouped", [("nccl", False), ("nvls", True)])
def test_static_shape_serving_dispatchers_match_training_dispatcher(kind, grouped):
    run_distributed(_serving_dispatch, 2, kind, grouped)


def _egtp(rank, world, grouped):
    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.distributed import DistributedDataParallel, DistributedDataParallelConfig
    from megatron_b200.core.tensor_parallel.gtp_api import apply_expert_gtp
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(expert_gtp_remat_size=2)
    model_parallel_cuda_manual_seed(1)
    assert ps.get_expert_gtp_weight_remat_world_size() == 2 and dist.get_world_size(ps.get_expert_data_parallel_group_without_gtp()) == world // 2

    def run(gtp):
        torch.manual_seed(5)
        model_parallel_cuda_manual_seed(1)                    # the expert-parallel RNG stream must restart too: both runs need the same weights
        cfg, mlp = _moe_layer(grouped=grouped)
        pre = apply_expert_gtp(mlp, prefetch=True) if gtp else None
        ddp = DistributedDataParallel(cfg, DistributedDataParallelConfig(overlap_grad_reduce=False, use_distributed_optimizer=False), mlp)
        x = torch.randn(6, 2, 32, generator=torch.Generator().manual_seed(50 + rank))
        for _ in range(2):
            y, _ = ddp(x)
            y.float().square().mean().backward()
        ddp.finish_grad_sync()
        return mlp, pre, y

    m0, _, y0 = run(False)
    m1, pre, y1 = run(True)
    assert torch.allclose(y0, y1, atol=1e-6)
    assert pre.stats["gathers"] > 0 and pre.stats["regathers"] > 0 and all(not h._gtp_outstanding for h in pre.modules)
    g0 = {n: p.main_grad.clone() for n, p in m0.named_parameters()}
    r = dist.get_rank(ps.get_expert_gtp_weight_remat_group())
    n_sharded = 0
    for n, p in m1.named_parameters():
        if n.endswith("_shard"):
            full = g0[n[: -len("_shard")]]
            k = full.shape[0] // 2
            assert p.shape[0] == k and getattr(p, "gtp_expert", False)
            assert torch.allclose(p.main_grad, full[r * k : (r + 1) * k], atol=1e-6, rtol=1e-4), n
            n_sharded += 1
        else:
            assert torch.allclose(p.main_grad, g0[n], atol=1e-6, rtol=1e-4), n
    assert n_sharded == (2 if grouped else 8)              # weight1/weight2 stacked, or fc1+fc2 of each of the 4 experts
    return True


@pytest.mark.parametrize("grouped", [True, False])
def test_expert_side_gtp_matches_plain_ddp(grouped):
    assert run_distributed(_egtp, 4, grouped) == [True] * 4
