"""Small runtime components: PP layer layouts, quantisation recipes, FP8 emulation, masked datasets, fault injection, restart wrapper,
hybrid layer allocation, MoE reference ops."""
import pytest
import torch


def test_pipeline_layer_layout_parsing():
    from megatron_b200.core.enums import LayerType
    from megatron_b200.core.transformer.pipeline_parallel_layer_layout import PipelineParallelLayerLayout as L

    lay = L.from_str("Et*3|(tt|)*2,m|L", 1)
    assert [len(s) for s in lay.flat] == [4, 2, 2, 1, 1]
    lay.validate_layer_layout(7, 1)
    lay = L.from_str("Ett|tt|tt|ttL", 2)  # 4 stages over pp=2 → 2 virtual chunks each
    assert lay.vpp == 2 and lay.get_layer_offset(vp_stage=1, pp_rank=1) == 6 and lay.get_num_layers_to_build(vp_stage=1, pp_rank=0) == 2
    assert lay.get_layer_id_list(LayerType.decoder, 0, 1) == [2, 3]
    with pytest.raises(ValueError):
        L.from_str("Ex|L", 1)


def test_quantization_recipe_first_match_wins():
    from megatron_b200.core.quantization import RecipeConfig

    r = RecipeConfig.from_dict({"configs": {"f8": {"recipe": "tensorwise", "fp8_format": "hybrid"}, "bf16": {"recipe": "none"}},
                                "matchers": [{"pattern": "decoder.layers.0.*", "config": "bf16"}, {"pattern": "*.linear_fc1", "config": "f8"}]})
    assert not r.match("decoder.layers.0.mlp.linear_fc1").enabled
    assert r.match("decoder.layers.3.mlp.linear_fc1").enabled and r.match("embedding") is None
    with pytest.raises(KeyError):
        RecipeConfig.from_dict({"configs": {}, "matchers": [{"pattern": "*", "config": "nope"}]})


def test_fp8_linear_emulation_forward_backward():
    from megatron_b200.core.fp8_utils import E4M3, Fp8LinearState, fp8_linear, quantize

    torch.manual_seed(0)
    x = torch.randn(4, 16, 64, requires_grad=True)
    w = (0.1 * torch.randn(32, 64)).requires_grad_(True)
    y = fp8_linear(x, w)
    y.sum().backward()
    ref = x.detach() @ w.detach().t()
    assert ((y.float() - ref).abs().max() / ref.abs().max()).item() < 0.08
    gx_ref = torch.ones(4, 16, 32) @ w.detach()
    assert ((x.grad - gx_ref).abs().max() / gx_ref.abs().max()).item() < 0.1 and w.grad.shape == w.shape
    q, inv = quantize(torch.tensor([1.0, -448.0, 3.0]), E4M3)
    assert torch.allclose(q.float() * inv, torch.tensor([1.0, -448.0, 3.0]), rtol=0.07)
    st = Fp8LinearState(history_len=4)
    for _ in range(3):
        fp8_linear(x, w, recipe="delayed", metas=st.metas)
    assert st.metas[0].amax_history.shape == (4,) and st.get_extra_state()[0]["scale"] is not None


def test_masked_datasets_are_deterministic_and_well_formed():
    from megatron_b200.core.datasets.masked_dataset import BERTMaskedDataset, MaskedDatasetConfig, T5MaskedDataset

    c = MaskedDatasetConfig(sequence_length=64, vocab_size=1000, sequence_length_decoder=32)
    b1, b2 = BERTMaskedDataset(c)[3], BERTMaskedDataset(c)[3]
    assert all(torch.equal(b1[k], b2[k]) for k in b1)
    assert b1["text"].shape == (64,) and 5 <= int(b1["loss_mask"].sum()) <= 12
    sel = b1["loss_mask"].bool()
    assert (b1["labels"][sel] != 0).all() and (b1["text"][0] == c.cls_id)
    t = T5MaskedDataset(c)[5]
    n_sent_enc = int((t["text_enc"] >= 1000 - c.num_sentinels).sum())
    n_sent_dec = int((t["text_dec"] >= 1000 - c.num_sentinels).sum())
    assert n_sent_enc == n_sent_dec >= 1 and t["text_dec"][0] == c.bos_id
    assert int(t["loss_mask"].sum()) == int(t["dec_mask"].sum())


def test_fault_injector_and_inprocess_restart():
    from megatron_b200.core.fault_injector import Fault, FaultInjector, FaultInjectorConfig, InjectedFaultError
    from megatron_b200.training.inprocess_restart import maybe_wrap_for_inprocess_restart

    inj = FaultInjector(FaultInjectorConfig(fault_type=Fault.WORKLOAD_EXC, ranks=[0], at_iteration=2), rank=0, world_size=1)
    inj.on_iteration(1)
    with pytest.raises(InjectedFaultError):
        inj.on_iteration(2)
    inj.on_iteration(3)  # fires once
    assert not FaultInjector(FaultInjectorConfig(ranks=[1]), rank=0, world_size=2).armed
    calls = {"n": 0}

    def flaky():
        calls["n"] += 1
        if calls["n"] < 3:
            raise RuntimeError("transient")
        return "done"

    assert maybe_wrap_for_inprocess_restart(flaky, max_restarts=3)() == "done" and calls["n"] == 3
    calls["n"] = 0
    with pytest.raises(RuntimeError):
        maybe_wrap_for_inprocess_restart(flaky, max_restarts=1)()


def test_hybrid_layer_allocation():
    from megatron_b200.core.ssm.mamba_hybrid_layer_allocation import allocate_layers

    assert allocate_layers(4, override_pattern="M*M-") == ["M", "*", "M", "-"]
    lay = allocate_layers(16, 0.25, 0.25)
    assert lay.count("*") == 4 and lay.count("-") == 4 and lay.count("M") == 8
    with pytest.raises(ValueError):
        allocate_layers(3, override_pattern="MM")


def test_moe_reference_ops_cpu():
    from megatron_b200 import ops

    torch.manual_seed(0)
    x = torch.randn(6, 8)
    out = ops.moe_gather_rows(x, torch.tensor([3, 1, 1, 5]), torch.tensor([1.0, 2.0, 3.0, 4.0]))
    assert torch.allclose(out[2], 3.0 * x[1])
    y = torch.randn(4, 8)
    pos, w = torch.tensor([[0, 2], [1, -1], [3, 0]]), torch.rand(3, 2)
    c = ops.moe_combine_rows(y, pos, w)
    assert torch.allclose(c[1], w[1, 0] * y[1], atol=1e-6) and torch.allclose(c[0], w[0, 0] * y[0] + w[0, 1] * y[2], atol=1e-6)
    ids, rmap, tpe, probs = ops.moe_topk_router(torch.randn(50, 8), 2, "softmax")
    assert rmap.sum(1).eq(2).all() and int(tpe.sum()) == 100 and torch.allclose(probs.sum(1), torch.ones(50), atol=1e-5)


def test_gated_delta_rule_chunked_matches_recurrence():
    from megatron_b200.core.ssm.gated_delta_net import gated_delta_rule_chunked, gated_delta_rule_recurrent

    torch.manual_seed(0)
    b, l, h, dk, dv = 2, 50, 3, 8, 6
    q = torch.nn.functional.normalize(torch.randn(b, l, h, dk), dim=-1)
    k = torch.nn.functional.normalize(torch.randn(b, l, h, dk), dim=-1)
    v, g, beta = torch.randn(b, l, h, dv), -torch.rand(b, l, h) * 0.3, torch.rand(b, l, h)
    o1, s1 = gated_delta_rule_recurrent(q, k, v, g, beta)
    o2, s2 = gated_delta_rule_chunked(q, k, v, g, beta, 16)
    assert (o1 - o2).abs().max().item() < 1e-4 and (s1 - s2).abs().max().item() < 1e-4
    # state carried across two calls == one long call
    oa, sa = gated_delta_rule_chunked(q[:, :32], k[:, :32], v[:, :32], g[:, :32], beta[:, :32], 16)
    ob, sb = gated_delta_rule_recurrent(q[:, 32:], k[:, 32:], v[:, 32:], g[:, 32:], beta[:, 32:], sa)
    assert (torch.cat([oa, ob], 1) - o1).abs().max().item() < 1e-4


def _hetero(rank, world):
    import json

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.heterogeneous import HeterogeneousTransformerConfig, get_gpt_heterogeneous_layer_spec
    from megatron_b200.core.transformer.identity_op import IdentityOp

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    blocks = {"block_configs": [{"attention": {"n_heads_in_group": 2}, "ffn": {"ffn_mult": 2.0}},
                                {"attention": {"no_op": True}, "ffn": {"ffn_hidden_size": 96}},
                                {"attention": {"replace_with_linear": True}, "ffn": {"no_op": True}}]}
    cfg = HeterogeneousTransformerConfig(num_layers=3, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0,
                                         normalization="RMSNorm", add_bias_linear=False, heterogeneous_layers_config_encoded_json=json.dumps(blocks))
    assert cfg.per_block_parameters[0].attention.num_query_groups == 2 and cfg.per_block_parameters[0].mlp.ffn_hidden_size == 256
    m = GPTModel(cfg, get_gpt_heterogeneous_layer_spec(cfg), vocab_size=128, max_sequence_length=32, position_embedding_type="rope")
    L = m.decoder.layers
    assert isinstance(L[1].self_attention, IdentityOp) and isinstance(L[2].mlp, IdentityOp) and L[1].mlp.linear_fc2.weight.shape[1] == 96
    tok = torch.randint(0, 128, (2, 16))
    m(tok, torch.arange(16)[None].expand(2, -1), None, labels=tok).mean().backward()
    assert all(p.grad is not None for p in m.parameters())
    return True


def test_heterogeneous_layer_specs():
    from dist_utils import run_distributed

    assert all(run_distributed(_hetero, 1))


def test_tp_comm_mode_resolution(monkeypatch):
    from megatron_b200.parallel import fused

    monkeypatch.setattr(fused, "_MODE", "auto")
    monkeypatch.setattr(fused, "_AUTO_FUSED_MIN_TP", 2)
    assert fused.get_mode(world_size=1) == "nccl" and fused.get_mode(world_size=2) == "fused" and fused.get_mode(world_size=8) == "fused"
    monkeypatch.setattr(fused, "_AUTO_FUSED_MIN_TP", 4)
    assert fused.get_mode(world_size=2) == "nccl"
    monkeypatch.setattr(fused, "_MODE", "nvlink")
    assert fused.get_mode(world_size=1) == "nvlink"
    # CPU tensors never take the NVLink path, whatever the mode
    assert fused._nvl(None, torch.zeros(1)) is None


def _gdn_hybrid(rank, world):
    import types

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.mamba.mamba_layer_specs import mamba_stack_spec
    from megatron_b200.core.models.mamba.mamba_model import MambaModel
    from megatron_b200.core.ssm.gated_delta_net import GatedDeltaNet, gated_delta_rule_chunked
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    # the reference's --linear-* geometry: 2 key heads x 16, 4 value heads x 16 (two value heads share a key head), conv kernel 4
    cfg = TransformerConfig(num_layers=4, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, is_hybrid_model=True,
                            linear_key_head_dim=16, linear_value_head_dim=16, linear_num_key_heads=2, linear_num_value_heads=4, linear_conv_kernel_dim=4)
    m = MambaModel(cfg, mamba_stack_spec, vocab_size=64, max_sequence_length=32, hybrid_override_pattern="G*G-")
    mixers = [l.mixer for l in m.decoder.layers if hasattr(l, "mixer")]
    assert [type(x) for x in mixers] == [GatedDeltaNet, GatedDeltaNet] and mixers[0].r == 2 and mixers[0].A_log.shape == (4,)
    assert mixers[0].in_proj.weight.shape == (2 * (2 * 16 + 2 * (2 * 16 + 2)), 64) and mixers[0].out_proj.weight.shape == (64, 64)
    tok = torch.randint(0, 64, (2, 32), generator=torch.Generator().manual_seed(0))
    pos = torch.arange(32).unsqueeze(0).expand(2, -1)
    loss = m(tok, pos, None, labels=tok).mean()
    loss.backward()
    # strong decays (A up to 16) used to overflow exp() above the diagonal of the in-chunk decay matrix: finite forward, NaN backward
    bad = [n for n, p in m.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not bad and float(mixers[0].in_proj.weight.grad.abs().sum()) > 0, bad
    q, k, v = torch.randn(1, 64, 2, 8), torch.randn(1, 64, 2, 8), torch.randn(1, 64, 2, 8, requires_grad=True)
    g = torch.full((1, 64, 2), -12.0, requires_grad=True)
    o, _ = gated_delta_rule_chunked(q, k, v, g, torch.rand(1, 64, 2), 64)
    o.sum().backward()
    assert torch.isfinite(g.grad).all() and torch.isfinite(v.grad).all()
    # prefill + one decode step equals the full forward (conv state and delta-rule state are carried in the inference context)
    mixer = mixers[0].eval()
    x = torch.randn(9, 2, 64)
    with torch.no_grad():
        full = mixer(x)[0]
        ctx = types.SimpleNamespace(sequence_len_offset=0, key_value_memory_dict={})
        head = mixer(x[:8], inference_context=ctx)[0]
        ctx.sequence_len_offset = 8
        last = mixer(x[8:], inference_context=ctx)[0]
    assert (torch.cat([head, last]) - full).abs().max().item() < 1e-4
    return True


def test_gated_delta_net_layers_in_the_hybrid_stack():
    from dist_utils import run_distributed

    assert run_distributed(_gdn_hybrid, 1) == [True]


def _hybrid_symbols(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.mamba.mamba_layer_specs import mamba_stack_spec
    from megatron_b200.core.models.mamba.mamba_model import MambaModel
    from megatron_b200.core.ssm.mamba_hybrid_layer_allocation import parse_hybrid_pattern
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import MLATransformerConfig

    assert parse_hybrid_pattern("M*M*|M-G+D/MM/MM") == (list("M*M*M-G+D"), [4, 5], ["MM", "MM"])
    assert parse_hybrid_pattern("M*-") == (list("M*-"), None, [])
    ps.initialize_model_parallel(1, world)                       # pipeline parallel over the ranks
    model_parallel_cuda_manual_seed(1)
    cfg = MLATransformerConfig(num_layers=5, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, is_hybrid_model=True,
                               q_lora_rank=32, kv_lora_rank=32, qk_head_dim=16, qk_pos_emb_head_dim=8, v_head_dim=16, rope_type="rope", add_bias_linear=False,
                               dsa_indexer_n_heads=2, dsa_indexer_head_dim=16, dsa_indexer_topk=8, pipeline_model_parallel_size=world, pipeline_dtype=torch.float32,
                               linear_key_head_dim=16, linear_value_head_dim=16, linear_num_key_heads=2, mamba_head_dim=16, mamba_num_groups=2, mamba_state_dim=16)
    # "|": stage 0 takes three layers, stage 1 two — an uneven split that num_layers // pp could not express
    m = MambaModel(cfg, mamba_stack_spec, vocab_size=64, max_sequence_length=32, hybrid_override_pattern="G+D|-M" if world == 2 else "G+D-M",
                   pre_process=ps.is_pipeline_first_stage(), post_process=ps.is_pipeline_last_stage(), position_embedding_type="rope")
    def kind(layer):
        for name in ("mixer", "self_attention", "mlp"):
            mod = getattr(layer, name, None)
            if mod is not None and type(mod).__name__ != "IdentityOp":
                return type(mod).__name__

    kinds = [kind(l) for l in m.decoder.layers]
    if world == 1:
        tok = torch.randint(0, 64, (2, 32), generator=torch.Generator().manual_seed(0))
        loss = m(tok, torch.arange(32).unsqueeze(0).expand(2, -1), None, labels=tok).mean()
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    return kinds


def test_hybrid_pattern_latent_attention_symbols_and_pipeline_separators():
    from dist_utils import run_distributed

    assert run_distributed(_hybrid_symbols, 1) == [["GatedDeltaNet", "MLASelfAttention", "DSAMLASelfAttention", "MLP", "MambaMixer"]]
    assert run_distributed(_hybrid_symbols, 2) == [["GatedDeltaNet", "MLASelfAttention", "DSAMLASelfAttention"], ["MLP", "MambaMixer"]]
