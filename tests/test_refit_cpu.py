"""Weight refit between parallel layouts (core/resharding: metadata -> run-intersection planner -> copy services -> refit API)."""
import itertools

import pytest
import torch

from dist_utils import run_distributed


# --------------------------------------------------------------------------------------------------------------------
# planner, in one process: simulate every rank, move bytes through a dictionary "wire"
# --------------------------------------------------------------------------------------------------------------------
def _shard(full, dim, world, rank, stride=1, sizes=None):
    """Canonical semantics of a TP shard: `stride` segments (or packed blocks of per-rank `sizes`), each chunked over `world`."""
    if sizes is not None:
        blocks = torch.split(full, [s * world for s in sizes], dim=dim)
    else:
        blocks = torch.chunk(full, stride, dim=dim)
    return torch.cat([torch.chunk(b, world, dim=dim)[rank] for b in blocks], dim=dim).contiguous()


def _simulate(metas_src, metas_dst, tensors_src, tensors_dst, world):
    """metas_*[rank] = [ParameterMetadata]; tensors_*[rank] = {name: tensor}.  Runs every rank's plan."""
    from megatron_b200.core.resharding.planner import build_plan_from_rosters

    gathered = [(metas_src.get(r, []), metas_dst.get(r, [])) for r in range(world)]
    plans = [build_plan_from_rosters(gathered, r) for r in range(world)]
    wire = {}
    for r, plan in enumerate(plans):
        for op in plan.send_ops:
            wire[op.task_id] = tensors_src[r][op.param_name][op.my_slice].clone()
    n_recv = 0
    for r, plan in enumerate(plans):
        for op in plan.recv_ops:
            tensors_dst[r][op.param_name][op.my_slice] = wire[op.task_id].to(tensors_dst[r][op.param_name].dtype)
            n_recv += 1
    assert n_recv == len(wire), "every send must be matched by exactly one receive"
    return plans


def _meta(name, t, rank, tp_ranks=None, dim=0, stride=1, sizes=None, ep_ranks=None, num_experts=None):
    from megatron_b200.core.resharding.utils import ParameterMetadata, assign_resolved_name_inplace

    m = ParameterMetadata(name=name, shape=tuple(t.shape), dtype=t.dtype, element_size=t.element_size(), is_tp=tp_ranks is not None and len(tp_ranks) > 1,
                          partition_dim=dim, partition_stride=stride, partition_sizes=sizes, is_ep=ep_ranks is not None, num_experts=num_experts, owner_rank=rank,
                          tensor_parallel_group_ranks=tp_ranks, expert_parallel_group_ranks=ep_ranks)
    assign_resolved_name_inplace(m, base_name=name, fused=(ep_ranks is not None and t.dim() == 3))
    return m


@pytest.mark.parametrize("src_tp,dst_tp,dim,stride", [(2, 4, 0, 1), (4, 2, 0, 2), (2, 1, 1, 1), (1, 4, 0, 2), (4, 4, 1, 1), (3, 2, 0, 1), (2, 3, 0, 2)])
def test_tp_refit_plan_moves_every_element_once(src_tp, dst_tp, dim, stride):
    world = max(src_tp, dst_tp)
    full = torch.arange(24 * 12, dtype=torch.float32).view(24, 12)
    src_ranks, dst_ranks = list(range(src_tp)), list(range(dst_tp))
    ts = {r: {"w": _shard(full, dim, src_tp, r, stride)} for r in src_ranks}
    td = {r: {"w": torch.full_like(_shard(full, dim, dst_tp, r, stride), -1.0)} for r in dst_ranks}
    ms = {r: [_meta("w", ts[r]["w"], r, src_ranks, dim, stride)] for r in src_ranks}
    md = {r: [_meta("w", td[r]["w"], r, dst_ranks, dim, stride)] for r in dst_ranks}
    plans = _simulate(ms, md, ts, td, world)
    for r in dst_ranks:
        assert torch.equal(td[r]["w"], _shard(full, dim, dst_tp, r, stride))
    if src_tp == dst_tp:
        assert all(op.peer_rank == r for r, p in enumerate(plans) for op in p.recv_ops), "identical layouts copy locally"
        assert all(len(p.recv_ops) == 1 for p in plans[:dst_tp]), "adjacent pieces are merged back into one transfer"


def test_block_interleaved_tp_refit_like_mamba_in_proj():
    # packed [z | x | B | C | dt] with different widths, each block sharded on its own
    full_sizes = [16, 16, 8, 8, 4]
    full = torch.randn(sum(full_sizes), 6)
    for src_tp, dst_tp in [(2, 4), (4, 1), (1, 2)]:
        s_sizes, d_sizes = [f // src_tp for f in full_sizes], [f // dst_tp for f in full_sizes]
        ts = {r: {"in_proj.weight": _shard(full, 0, src_tp, r, sizes=s_sizes)} for r in range(src_tp)}
        td = {r: {"in_proj.weight": torch.zeros_like(_shard(full, 0, dst_tp, r, sizes=d_sizes))} for r in range(dst_tp)}
        ms = {r: [_meta("in_proj.weight", ts[r]["in_proj.weight"], r, list(range(src_tp)), 0, sizes=s_sizes)] for r in range(src_tp)}
        md = {r: [_meta("in_proj.weight", td[r]["in_proj.weight"], r, list(range(dst_tp)), 0, sizes=d_sizes)] for r in range(dst_tp)}
        _simulate(ms, md, ts, td, max(src_tp, dst_tp))
        for r in range(dst_tp):
            assert torch.equal(td[r]["in_proj.weight"], _shard(full, 0, dst_tp, r, sizes=d_sizes))


def test_expert_parallel_refit_fused_and_per_expert_names():
    E, F, H = 8, 6, 4
    full = torch.randn(E, F, H)
    # fused grouped tensor: EP=4 (2 local experts) x expert-TP=1  ->  EP=2 (4 local) x TP... expert axis + TP on dim 1
    src_ep, dst_ep = 4, 2
    ts = {r: {"experts.weight1": full.chunk(src_ep, 0)[r].clone()} for r in range(src_ep)}
    td = {r: {"experts.weight1": torch.zeros(E // dst_ep, F, H)} for r in range(dst_ep)}
    ms = {r: [_meta("experts.weight1", ts[r]["experts.weight1"], r, None, ep_ranks=list(range(src_ep)), num_experts=E)] for r in range(src_ep)}
    md = {r: [_meta("experts.weight1", td[r]["experts.weight1"], r, None, ep_ranks=list(range(dst_ep)), num_experts=E)] for r in range(dst_ep)}
    _simulate(ms, md, ts, td, 4)
    for r in range(dst_ep):
        assert torch.equal(td[r]["experts.weight1"], full.chunk(dst_ep, 0)[r])
    # per-expert modules: local_experts.{i} is a LOCAL index -> matched through the global expert index
    ts = {r: {f"experts.local_experts.{i}.linear_fc2.weight": full[r * 2 + i].clone() for i in range(2)} for r in range(4)}
    td = {r: {f"experts.local_experts.{i}.linear_fc2.weight": torch.zeros(F, H) for i in range(4)} for r in range(2)}
    ms = {r: [_meta(n, t, r, None, ep_ranks=[0, 1, 2, 3], num_experts=E) for n, t in ts[r].items()] for r in range(4)}
    md = {r: [_meta(n, t, r, None, ep_ranks=[0, 1], num_experts=E) for n, t in td[r].items()] for r in range(2)}
    assert ms[3][1].resolved_name == "experts.local_experts.7.linear_fc2.weight" and ms[3][1].global_expert_index == 7
    _simulate(ms, md, ts, td, 4)
    for r in range(2):
        for i in range(4):
            assert torch.equal(td[r][f"experts.local_experts.{i}.linear_fc2.weight"], full[r * 4 + i])


def test_replicated_sources_are_balanced_and_local_first():
    full = torch.randn(8, 4)
    # two DP replicas of a TP=1 source (ranks 0, 1); four TP=4 destinations (ranks 0..3)
    ts = {r: {"w": full.clone()} for r in (0, 1)}
    td = {r: {"w": torch.zeros(2, 4)} for r in range(4)}
    ms = {r: [_meta("w", full, r)] for r in (0, 1)}
    md = {r: [_meta("w", td[r]["w"], r, [0, 1, 2, 3], 0)] for r in range(4)}
    plans = _simulate(ms, md, ts, td, 4)
    src_of = {r: plans[r].recv_ops[0].peer_rank for r in range(4)}
    assert src_of[0] == 0 and src_of[1] == 1, "a rank that holds a replica reads its own"
    assert {src_of[2], src_of[3]} == {0, 1}, "remote destinations are spread over the replicas"
    for r in range(4):
        assert torch.equal(td[r]["w"], full.chunk(4, 0)[r])


def test_missing_and_mismatched_sources_are_reported():
    from megatron_b200.core.resharding.planner import build_plan_from_rosters

    a = torch.zeros(4, 4)
    with pytest.raises(KeyError):
        build_plan_from_rosters([([_meta("x", a, 0)], [_meta("y", a, 0)])], 0)
    with pytest.raises(ValueError):
        build_plan_from_rosters([([_meta("x", a, 0)], [_meta("x", torch.zeros(4, 5), 0)])], 0)
    with pytest.raises(ValueError):      # TP=2 source with only rank 0 present: half of the tensor has no holder
        build_plan_from_rosters([([_meta("x", torch.zeros(2, 4), 0, [0, 1], 0)], [_meta("x", a, 0)])], 0)


def test_pipeline_local_layer_numbers_resolve_to_global_names():
    from megatron_b200.core.resharding.utils import extract_module_metadata

    class Layer(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            self.layer_number = n                        # 1-based GLOBAL number, as TransformerLayer keeps it
            self.w = torch.nn.Parameter(torch.zeros(2))
            self.register_buffer("expert_bias", torch.zeros(2))
            self.register_buffer("scratch", torch.zeros(2), persistent=False)

    class Stage(torch.nn.Module):
        def __init__(self, first):
            super().__init__()
            self.layers = torch.nn.ModuleList([Layer(first + i + 1) for i in range(2)])

    metas = extract_module_metadata(Stage(2), owner_rank=1)          # second pipeline stage: local 0, 1 = global 2, 3
    names = {m.name: m.resolved_name for m in metas}
    assert names["layers.0.w"] == "layers.2.w" and names["layers.1.expert_bias"] == "layers.3.expert_bias"
    assert not any("scratch" in n for n in names), "non-persistent buffers are not part of a refit"


# --------------------------------------------------------------------------------------------------------------------
# copy services + refit API over gloo
# --------------------------------------------------------------------------------------------------------------------
class _Model(torch.nn.Module):
    """decoder.layers.{i}.mlp with a gated fc1 — the smallest thing with strided TP shards, plain TP shards, replicated
    parameters and a persistent buffer."""

    def __init__(self, config, tp_group, layer_numbers):
        super().__init__()
        from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
        from megatron_b200.core.transformer.mlp import MLP

        self.config = config
        mlp_spec = get_gpt_layer_local_spec().submodules.mlp
        layers = []
        for n in layer_numbers:
            layer = torch.nn.Module()
            layer.layer_number = n
            layer.mlp = MLP(config, mlp_spec.submodules, tp_group=tp_group)
            layer.norm_weight = torch.nn.Parameter(torch.randn(config.hidden_size))
            layer.register_buffer("steps", torch.zeros(1))
            layers.append(layer)
        self.layers = torch.nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            y, _ = layer.mlp(x * layer.norm_weight)
            x = x + y
        return x


class _Groups:
    def __init__(self, **kw):
        self.tp = self.pp = self.dp = self.ep = self.expt_tp = self.expt_dp = None
        self.__dict__.update(kw)


def _refit_worker(rank, world, backend_kind):
    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.resharding import clear_all_caches, prepare_swap_model_weights, swap_model_weights
    from megatron_b200.core.resharding.refit import _PLANS
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel(tensor_model_parallel_size=2)
    cfg = TransformerConfig(num_layers=2, hidden_size=16, num_attention_heads=4, ffn_hidden_size=24, gated_linear_unit=True, activation_func=torch.nn.functional.silu,
                            add_bias_linear=True, use_cpu_initialization=True, bias_activation_fusion=False)
    tp2 = ps.get_tensor_model_parallel_group()
    dp2 = ps.get_data_parallel_group()
    tp4 = dist.group.WORLD
    torch.manual_seed(1234 + dist.get_rank(dp2) * 0)       # DP replicas hold identical weights
    src = _Model(cfg, tp2, [1, 2])
    src.pg_collection = _Groups(tp=tp2, dp=dp2)
    for layer in src.layers:                               # make replicated params identical across the whole world, shards distinct per TP rank
        for p in layer.mlp.parameters():
            torch.manual_seed(7 + dist.get_rank(tp2) * 100 + p.numel())
            p.data.copy_(0.2 * torch.randn_like(p))
        torch.manual_seed(99)
        layer.norm_weight.data.copy_(torch.randn(16))
        layer.steps.fill_(5.0)
        fc2b = layer.mlp.linear_fc2.bias
        torch.manual_seed(3)
        fc2b.data.copy_(torch.randn_like(fc2b))            # replicated bias of the row-parallel linear
    dst = _Model(cfg, tp4, [1, 2])
    dst.pg_collection = _Groups(tp=tp4)
    for p in dst.parameters():
        p.data.fill_(float("nan"))
    plan = prepare_swap_model_weights(src, dst, group=dist.group.WORLD)
    assert len(_PLANS) == 1
    swap_model_weights(src, dst, backend_kind, group=dist.group.WORLD)
    assert len(_PLANS) == 1, "the second call reuses the cached plan"
    assert not any(torch.isnan(p).any() for p in dst.parameters())
    assert float(dst.layers[1].steps) == 5.0
    # same function: TP=2 source and TP=4 destination give the same output on the same input
    torch.manual_seed(0)
    x = torch.randn(3, 2, 16)
    with torch.no_grad():
        y_src, y_dst = src(x), dst(x)
    assert torch.allclose(y_src, y_dst, atol=1e-4, rtol=1e-4), float((y_src - y_dst).abs().max())
    # and back into a fresh TP=2 model: bit-identical shards
    back = _Model(cfg, tp2, [1, 2])
    back.pg_collection = _Groups(tp=tp2, dp=dp2)
    swap_model_weights(dst, back, backend_kind, group=dist.group.WORLD)
    for (n, a), (_, b) in zip(src.named_parameters(), back.named_parameters()):
        assert torch.equal(a, b), n
    sends = sum(op.nbytes for op in plan.send_ops if op.peer_rank != rank)
    clear_all_caches()
    return sends


@pytest.mark.parametrize("backend_kind", ["gloo"])
def test_swap_model_weights_tp2dp2_to_tp4_and_back(backend_kind):
    out = run_distributed(_refit_worker, 4, backend_kind)
    assert all(v >= 0 for v in out.values() if v is not None) if isinstance(out, dict) else True


def _non_collocated_worker(rank, world):
    """Ranks 0-1 train (TP=2), ranks 2-3 serve (TP=2 as well, but they only RECEIVE); a bf16 trainer feeds an fp32 server."""
    import torch.distributed as dist

    from megatron_b200.core.resharding import clear_all_caches, swap_model_weights
    from megatron_b200.core.tensor_parallel.layers import set_tensor_model_parallel_attributes

    class M(torch.nn.Module):
        def __init__(self, dtype):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(4, 6, dtype=dtype))
            set_tensor_model_parallel_attributes(self.w, True, 0, 1)
            self.b = torch.nn.Parameter(torch.zeros(6, dtype=dtype))

    full = torch.arange(48, dtype=torch.float32).view(8, 6)
    if rank < 2:
        m = M(torch.bfloat16)
        m.w.data.copy_(full.chunk(2, 0)[rank])
        m.b.data.copy_(torch.arange(6.0))
        m.pg_collection = _Groups(tp=[0, 1])               # group-local ranks of the trainer world
        swap_model_weights(m, None, "gloo", group=dist.group.WORLD, src_rank_offset=0, dst_rank_offset=2)
    else:
        m = M(torch.float32)
        m.pg_collection = _Groups(tp=[0, 1])               # group-local ranks of the server world, shifted by dst_rank_offset=2
        swap_model_weights(None, m, "gloo", group=dist.group.WORLD, src_rank_offset=0, dst_rank_offset=2)
        assert torch.equal(m.w.data, full.chunk(2, 0)[rank - 2]) and torch.equal(m.b.data, torch.arange(6.0))
    clear_all_caches()
    return True


def test_non_collocated_refit_with_rank_offsets_and_dtype_change():
    run_distributed(_non_collocated_worker, 4)


def test_mxfp8_transform_quantises_on_receive():
    from megatron_b200.core.resharding.copy_services.base import CopyService
    from megatron_b200.core.resharding.execution import execute_reshard_plan
    from megatron_b200.core.resharding.planner import build_plan_from_rosters
    from megatron_b200.core.resharding.transforms import MXFP8ReshardTransform
    from megatron_b200.core.resharding.utils import extract_module_metadata
    from megatron_b200.ops.extra import mxfp8_dequantize, mxfp8_quantize_reference

    class Local(CopyService):
        def run(self):
            _, _, local = self._take()
            for s, r in local:
                r.tensor.copy_(s.tensor)

    src, dst = torch.nn.Linear(64, 8, bias=False).bfloat16(), torch.nn.Linear(64, 8, bias=False).bfloat16()
    payload, scales = torch.zeros(8, 64, dtype=torch.uint8), torch.zeros(8, 2, dtype=torch.uint8)
    refreshed = []
    tr = MXFP8ReshardTransform({"weight": (payload, scales)}, refresh=refreshed.append)
    plan = build_plan_from_rosters([(extract_module_metadata(src, 0), extract_module_metadata(dst, 0))], 0)
    before = dst.weight.detach().clone()
    execute_reshard_plan(plan, src, dst, Local(), transform=tr)
    q, sf = mxfp8_quantize_reference(src.weight.detach())
    assert torch.equal(payload, q) and torch.equal(scales, sf) and refreshed == ["weight"]
    assert torch.equal(dst.weight, before), "the bf16 destination parameter is not written when the transform claims the tensor"
    assert (mxfp8_dequantize(payload, scales).float() - src.weight.float()).abs().max() < 0.08
