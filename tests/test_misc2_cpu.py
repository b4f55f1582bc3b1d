"""Inference TP layers, RADIO tower, Hugging Face wrappers, dataset merge tool."""
import os
import subprocess
import sys


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]

import numpy as np
import torch

from dist_utils import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inference_layers(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.inference_layers import InferenceColumnParallelLinear, InferenceRowParallelLinear, convert_to_inference_layers
    from megatron_b200.core.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.torch_norm import FusedNorm
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    model_parallel_cuda_manual_seed(1)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True, tensor_model_parallel_size=world, sequence_parallel=True,
                            normalization="RMSNorm", add_bias_linear=False)
    torch.manual_seed(3)
    col = ColumnParallelLinear(32, 64, config=cfg, init_method=cfg.init_method, bias=False, gather_output=False)
    row = RowParallelLinear(64, 32, config=cfg, init_method=cfg.init_method, bias=False, input_is_parallel=True, skip_bias_add=True)
    norm = FusedNorm(cfg, 32, eps=1e-5)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
    torch.manual_seed(10 + rank)
    x_shard, res_shard = torch.randn(8 // world, 2, 32), torch.randn(8 // world, 2, 32)
    # training layers (autograd mappings) as the oracle
    h, _ = col(x_shard)
    y, _ = row(torch.tanh(h))
    new_res = res_shard + y
    normed = norm(new_res)
    full = torch.cat([torch.empty_like(normed) for _ in range(world)])
    torch.distributed.all_gather_into_tensor(full, normed.detach().contiguous(), group=ps.get_tensor_model_parallel_group())
    mod = torch.nn.ModuleList([col, row])
    convert_to_inference_layers(mod)
    assert isinstance(col, InferenceColumnParallelLinear) and isinstance(row, InferenceRowParallelLinear)
    h2, _ = col(x_shard)
    normed2, res2 = row(torch.tanh(h2), residual=res_shard, norm=norm)
    assert not h2.requires_grad and torch.allclose(h2, h.detach(), atol=1e-6)
    assert torch.allclose(res2, new_res.detach(), atol=1e-5) and torch.allclose(normed2, full, atol=1e-5) and normed2.shape == (8, 2, 32)
    out_plain, _ = row(torch.tanh(h2))
    assert torch.allclose(out_plain, y.detach(), atol=1e-5)
    return True


def test_inference_tp_layers_match_training_layers():
    assert run_distributed(_inference_layers, 2) == [True, True]


def _towers(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.huggingface import build_hf_model
    from megatron_b200.core.models.vision.radio import RADIOViTModel
    from megatron_b200.core.models.vision.vit_layer_specs import get_vit_layer_with_local_spec
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
    m = RADIOViTModel(cfg, get_vit_layer_with_local_spec(), patch_dim=4, img_h=16, img_w=16, max_img_h=32, max_img_w=32, class_token_len=2, num_registers=3)
    out = m(torch.randn(2, 3, 16, 16))
    assert out.shape == (2, 2 + 16, 32)
    out2 = m(torch.randn(1, 3, 32, 24))                  # another resolution through the interpolated position table
    assert out2.shape == (1, 2 + 8 * 6, 32)
    out.sum().backward()
    assert m.position_embeddings.grad is not None and m.embedder.weight.grad is not None
    import transformers

    hf_cfg = transformers.BertConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64, vocab_size=50, max_position_embeddings=16)
    w = build_hf_model(cfg, hf_config=hf_cfg, model_cls="BertModel")
    hs = w(input_ids=torch.randint(0, 50, (2, 8)))
    assert hs.shape == (2, 8, 32) and all(hasattr(p, "sequence_parallel") for p in w.parameters())
    return True


def test_radio_tower_and_hf_wrapper():
    assert run_distributed(_towers, 1) == [True]


def test_merge_datasets_tool(tmp_path):
    from megatron_b200.core.datasets.indexed_dataset import IndexedDataset, IndexedDatasetBuilder, get_bin_path, get_idx_path

    docs = {"a": [[1, 2, 3], [4, 5]], "b": [[6], [7, 8, 9, 10]]}
    for name, dd in docs.items():
        b = IndexedDatasetBuilder(get_bin_path(str(tmp_path / name)), dtype=np.int32)
        for d in dd:
            b.add_document(torch.tensor(d, dtype=torch.int32), [len(d)])
        b.finalize(get_idx_path(str(tmp_path / name)))
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "merge_datasets.py"), "--input", str(tmp_path), "--output-prefix", str(out / "merged")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    ds = IndexedDataset(str(out / "merged"))
    assert [ds[i].tolist() for i in range(len(ds))] == docs["a"] + docs["b"]


def _retro(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.retro import RetroConfig, RetroModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    cfg = RetroConfig(num_layers=3, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0,
                      retro_chunk_length=4, retro_num_neighbors=2, retro_retrieved_length=6, retro_encoder_num_layers=1, retro_decoder_cross_attention_layers=[2, 3],
                      retro_encoder_hidden_dropout=0.0, retro_encoder_attention_dropout=0.0)
    torch.manual_seed(2)
    m = RetroModel(cfg, get_gpt_layer_local_spec(), vocab_size=64, max_sequence_length=32, position_embedding_type="learned_absolute")
    b, n, l = 2, 16, 4
    ids = torch.randint(0, 64, (b, n))
    pos = torch.arange(n)[None].expand(b, -1)
    ctx = torch.randint(0, 64, (b, l, 2, 6))
    loss = m(ids, pos, None, context_input_ids=ctx, labels=ids).float().mean()
    loss.backward()
    missing = [n_ for n_, p in m.named_parameters() if p.grad is None]
    assert not missing, missing
    # autoregressive retrieval: the neighbours of chunk u are first visible to the LAST token of chunk u (position u·m + m - 1)
    m.eval()
    with torch.no_grad():
        base = m(ids, pos, None, context_input_ids=ctx)
        ctx2 = ctx.clone()
        ctx2[:, 2] = torch.randint(0, 64, (b, 2, 6))            # change the neighbours of chunk 2 (tokens 8..11)
        alt = m(ids, pos, None, context_input_ids=ctx2)
        first_visible = 2 * 4 + 3
        assert torch.allclose(base[:, :first_visible], alt[:, :first_visible], atol=1e-6)
        assert (base[:, first_visible:] - alt[:, first_visible:]).abs().max() > 1e-5
        assert not torch.allclose(m(ids, pos, None), base)      # retrieval changes the prediction at all
    return True


def test_retro_chunked_cross_attention_is_causal_in_the_neighbours():
    assert run_distributed(_retro, 1) == [True]


def _mixtral_export(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.export.hf_mixtral import hf_mixtral_to_megatron, hub_to_fused_experts_layout, megatron_to_hf_mixtral
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_decoder_block_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    out = []
    for grouped in (True, False):
        cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=48, use_cpu_initialization=True, normalization="RMSNorm",
                                gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False, hidden_dropout=0.0, attention_dropout=0.0, num_moe_experts=4,
                                moe_router_topk=2, moe_grouped_gemm=grouped, moe_token_dispatcher_type="alltoall", moe_aux_loss_coeff=0.0)
        torch.manual_seed(4)
        m = GPTModel(cfg, get_gpt_decoder_block_spec(cfg), vocab_size=64, max_sequence_length=16, position_embedding_type="rope", share_embeddings_and_output_weights=False)
        sd = {k: v for k, v in m.state_dict().items() if isinstance(v, torch.Tensor) and "_extra_state" not in k}
        hf = megatron_to_hf_mixtral(sd, 4, 2, 8)
        assert hf["model.layers.1.block_sparse_moe.experts.3.w2.weight"].shape == (32, 48) and hf["model.layers.0.block_sparse_moe.gate.weight"].shape == (4, 32)
        back = hf_mixtral_to_megatron(hf, 4, 2, 8, grouped=grouped)
        assert set(back) == set(sd), (set(back) ^ set(sd))
        assert all(torch.equal(back[k], sd[k]) for k in sd)
        # the HF-layout weights really are a Mixtral: load them into transformers' implementation and compare logits
        import transformers

        hc = transformers.MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                        num_local_experts=4, num_experts_per_tok=2, max_position_embeddings=16, rms_norm_eps=cfg.layernorm_epsilon, rope_theta=10000.0,
                                        tie_word_embeddings=False, attention_dropout=0.0, router_jitter_noise=0.0)
        hm = transformers.MixtralForCausalLM(hc).eval()
        sd_hf = hub_to_fused_experts_layout(hf) if any(k.endswith("experts.gate_up_proj") for k in hm.state_dict()) else hf
        missing, unexpected = hm.load_state_dict(sd_hf, strict=False)
        assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
        tok = torch.randint(0, 64, (2, 16))
        with torch.no_grad():
            ours = m.eval()(tok, torch.arange(16)[None].expand(2, -1), None)
            theirs = hm(tok).logits
        out.append((ours - theirs).abs().max().item())
    return out


def test_hf_mixtral_roundtrip_and_logits_match_transformers():
    (errs,) = run_distributed(_mixtral_export, 1)
    assert max(errs) < 2e-4, errs


def test_every_module_imports_and_entry_scripts_compile():
    """Reference ``tests/unit_tests/test_imports.py``: no module may fail at import time (missing optional deps must be gated)."""
    import importlib
    import pkgutil
    import py_compile

    import megatron_b200

    failed = []
    for m in pkgutil.walk_packages(megatron_b200.__path__, "megatron_b200."):
        if m.name.endswith("._C") or "helpers_cpp" in m.name:
            continue
        try:
            importlib.import_module(m.name)
        except Exception as e:  # noqa: BLE001
            failed.append((m.name, f"{type(e).__name__}: {e}"))
    assert not failed, failed
    for f in ("pretrain_gpt.py", "pretrain_bert.py", "pretrain_t5.py", "pretrain_mamba.py", "pretrain_hybrid.py", "pretrain_vlm.py", "train_rl.py", "bench.py", "gpt_builders.py",
              "model_provider.py", "setup.py", "__graft_entry__.py", "examples/run_simple_mcore_train_loop.py", "tools/run_dynamic_text_generation_server.py", "tools/merge_datasets.py"):
        py_compile.compile(os.path.join(ROOT, f), doraise=True)


def test_simple_mcore_train_loop_example_tp2():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "examples", "run_simple_mcore_train_loop.py"), "--tp", "2", "--iters", "3"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and "checkpoint round trip ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def _rl_infra(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.rl.agent import RewardOnlyAgent, Rollout, RolloutBank, WeightedMultiAgent
    from megatron_b200.rl.inference_interface import LocalEngineInference, WeightRefitter

    ps.initialize_model_parallel()

    def build(seed):
        torch.manual_seed(seed)
        cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=2, ffn_hidden_size=64, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                                normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
        return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=32, max_sequence_length=64, position_embedding_type="rope")

    policy, gen_model = build(1), build(2)
    even = RewardOnlyAgent(lambda n: [[1, 2, 3]] * n, lambda p, c: float(sum(t % 2 == 0 for t in c)))
    odd = RewardOnlyAgent(lambda n: [[4, 5]] * n, lambda p, c: float(sum(t % 2 == 1 for t in c)))
    agent = WeightedMultiAgent([even, odd], [0.5, 0.5], seed=3)
    sp = SamplingParams(temperature=1.0, num_tokens_to_generate=5, seed=0)
    for dynamic in (False, True):
        inf = LocalEngineInference(gen_model, vocab_size=32, dynamic=dynamic, num_blocks=64, block_size=4) if dynamic else LocalEngineInference(gen_model, vocab_size=32)
        inf.set_policy_version(7)
        groups = agent.rollout_groups(inf, n_prompts=3, group_size=4, sampling=sp)
        assert len(groups) == 3 and all(len(g) == 4 for g in groups)
        for g in groups:
            for r in g:
                assert len(r.completion) == 5 and r.policy_version == 7
                want = sum(t % 2 == (0 if r.prompt == [1, 2, 3] else 1) for t in r.completion)
                assert r.reward == float(want)
        assert gen_model.training       # generation restores the mode
    # refit: generation model takes the policy's weights
    rf = WeightRefitter(policy, gen_model)
    assert rf.refit() == 1 and all(torch.equal(a, b) for a, b in zip(policy.parameters(), gen_model.parameters()))
    # bank: staleness and uninformative groups
    bank = RolloutBank(max_staleness=1)
    mk = lambda v, rewards: [Rollout([1], [2], r, v) for r in rewards]  # noqa: E731
    bank.add([mk(3, [0, 1]), mk(5, [1, 1]), mk(5, [0, 2]), mk(6, [3, 1]), mk(6, [0, 1])])
    got = bank.sample(2, current_version=6)
    assert [min(r.policy_version for r in g) for g in got] == [5, 6] and bank.dropped_stale == 1           # version 3 is too old
    assert len(bank) == 2                                                                                  # the all-equal group and the unused fresh one stay
    return True


def test_rl_agents_inference_interface_refit_and_rollout_bank():
    assert run_distributed(_rl_infra, 1) == [True]


def _elastic(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.elastification import BudgetSampler, ElasticBudget, ElasticController, distillation_step, extract_mlp_subnetwork

    ps.initialize_model_parallel()
    torch.manual_seed(2)
    cfg = TransformerConfig(num_layers=3, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                            normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
    m = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=48, max_sequence_length=16, position_embedding_type="rope")
    tok = torch.randint(0, 48, (2, 16))
    pos = torch.arange(16)[None].expand(2, -1)
    fwd = lambda mod, b: mod(b, pos, None)  # noqa: E731
    base = fwd(m, tok).detach()
    ctl = ElasticController(m)
    assert torch.equal(fwd(m, tok), base)                                   # full budget = the original model, bit for bit
    ctl.calibrate([tok, torch.randint(0, 48, (2, 16))], fwd)
    assert sorted(m.decoder.layers[0].elastic_ffn_rank.tolist()) == list(range(64)) and sorted(ctl.layer_rank.tolist()) == [0, 1, 2]
    assert torch.equal(fwd(m, tok), base)                                   # ranking alone changes nothing
    # nestedness: the half-width network's active units are a subset of the three-quarter one's
    r = m.decoder.layers[1].elastic_ffn_rank
    assert set(torch.nonzero(r < 32).flatten().tolist()) <= set(torch.nonzero(r < 48).flatten().tolist())
    ctl.set_budget(ElasticBudget(ffn_fraction=0.5, head_fraction=0.5))
    half = fwd(m, tok).detach()
    assert not torch.allclose(half, base) and 0.45 < ctl.active_parameter_fraction() < 0.55
    # the masked MLP equals a physically sliced one
    L = m.decoder.layers[0]
    sub = extract_mlp_subnetwork(L, 0.5)
    assert sub["linear_fc1.weight"].shape == (64, 32) and sub["linear_fc2.weight"].shape == (32, 32)
    x = torch.randn(5, 2, 32)
    masked, _ = L.mlp(x)
    g, u = (x @ sub["linear_fc1.weight"].t()).chunk(2, -1)
    assert torch.allclose(masked, (F.silu(g) * u) @ sub["linear_fc2.weight"].t(), atol=1e-5)
    # importance ordering is useful: keeping the top half hurts less than keeping the bottom half
    err_top = (half - base).pow(2).mean()
    for Lx in m.decoder.layers:
        Lx.elastic_ffn_rank.copy_(63 - Lx.elastic_ffn_rank)
        Lx.elastic_head_rank.copy_(3 - Lx.elastic_head_rank)
    err_bottom = (fwd(m, tok).detach() - base).pow(2).mean()
    for Lx in m.decoder.layers:
        Lx.elastic_ffn_rank.copy_(63 - Lx.elastic_ffn_rank)
        Lx.elastic_head_rank.copy_(3 - Lx.elastic_head_rank)
    assert err_top < err_bottom
    # layer dropping and one sandwich-rule distillation step
    ctl.set_budget(ElasticBudget(layer_fraction=2 / 3))
    assert not torch.allclose(fwd(m, tok), base)
    loss = distillation_step(ctl, BudgetSampler(ffn=(0.5, 1.0), heads=(0.5, 1.0), n_random=1), tok, fwd,
                             lambda lg, b: F.cross_entropy(lg.reshape(-1, lg.shape[-1]).float(), b.reshape(-1)))
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())
    ctl.remove()
    assert torch.equal(fwd(m, tok), base)
    return True


def test_elastic_nested_subnetworks():
    assert run_distributed(_elastic, 1) == [True]


def _ptq(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.post_training.quantize import PTQConfig, calibrate, export_quantized_state_dict, quantize_model

    ps.initialize_model_parallel()
    tok = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16)[None].expand(2, -1)
    fwd = lambda m, b: m(b, pos, None)  # noqa: E731
    errs = {}
    for fmt in ("fp8", "mxfp8", "nvfp4", "w4a16"):
        torch.manual_seed(3)
        cfg = TransformerConfig(num_layers=2, hidden_size=256, num_attention_heads=4, ffn_hidden_size=512, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                                normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
        m = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=16, position_embedding_type="rope").eval()
        with torch.no_grad():
            base = fwd(m, tok)
        pcfg = PTQConfig(default=fmt, matchers=[("*output_layer*", "none"), ("decoder.layers.0.self_attention.linear_qkv", "none")])
        amax = calibrate(m, [tok], fwd, pcfg)
        assert "decoder.layers.1.mlp.linear_fc1" in amax and "output_layer" not in amax
        before = sum(p.numel() * p.element_size() for p in m.parameters())
        states = quantize_model(m, pcfg, amax)
        assert len(states) == 7 and "decoder.layers.0.self_attention.linear_qkv" not in states            # 2 x (qkv, proj, fc1, fc2) minus the excluded qkv
        after = sum(p.numel() * p.element_size() for p in m.parameters()) + sum(s.nbytes() for s in states.values())
        assert after < before * (0.62 if fmt in ("fp8", "mxfp8") else 0.5)
        with torch.no_grad():
            q = fwd(m, tok)
        errs[fmt] = ((q - base).norm() / base.norm()).item()
        sd = export_quantized_state_dict(m, states)
        assert "decoder.layers.1.mlp.linear_fc2.weight_q" in sd and "decoder.layers.1.mlp.linear_fc2.weight" not in sd and "output_layer.weight" in sd
    assert errs["fp8"] < 0.08 and errs["mxfp8"] < 0.08 and errs["w4a16"] < 0.25 and errs["nvfp4"] < 0.35 and errs["w4a16"] <= errs["nvfp4"] + 1e-6, errs
    return True


def test_post_training_quantisation_formats():
    assert run_distributed(_ptq, 1) == [True]
