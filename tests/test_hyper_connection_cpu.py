"""Manifold-constrained hyper-connections: Sinkhorn projection (values + gradient), the mappings' ranges, the layer formula, and a GPT model that trains with an
n-wide residual stream (reference tests/unit_tests/transformer/test_hyper_connection.py)."""
import torch
import torch.nn.functional as F

from dist_utils import run_distributed


def test_sinkhorn_projection_and_gradient():
    from megatron_b200.core.transformer.hyper_connection import SinkhornKnopp, sinkhorn_normalize

    torch.manual_seed(0)
    logits = torch.randn(3, 2, 4, 4, dtype=torch.float64, requires_grad=True)
    m = SinkhornKnopp.apply(logits, 30)
    assert torch.allclose(m.sum(-1), torch.ones(3, 2, 4, dtype=torch.float64), atol=1e-6) and torch.allclose(m.sum(-2), torch.ones(3, 2, 4, dtype=torch.float64), atol=1e-6)
    assert (m > 0).all()
    w = torch.randn_like(m)
    (g,) = torch.autograd.grad((m * w).sum(), logits)
    l2 = logits.detach().clone().requires_grad_(True)
    (g2,) = torch.autograd.grad((sinkhorn_normalize(torch.exp(l2), 30) * w).sum(), l2)          # plain autograd through the iterations, no shift
    assert torch.allclose(g, g2, atol=1e-8)
    assert torch.autograd.gradcheck(lambda x: SinkhornKnopp.apply(x, 5), (torch.randn(1, 1, 3, 3, dtype=torch.float64, requires_grad=True),), atol=1e-6)
    big = torch.full((1, 1, 2, 2), 500.0)                                                        # exp would overflow without the row-max shift
    assert torch.isfinite(SinkhornKnopp.apply(big, 3)).all()


def _mhc_worker(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.hyper_connection import HyperConnectionModule
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(3)
    torch.manual_seed(3)
    kw = dict(num_layers=2, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False, normalization="RMSNorm",
              use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, bias_dropout_fusion=False)
    cfg = TransformerConfig(enable_mhc_connections=True, mhc_num_residual_streams=3, mhc_sinkhorn_iterations=10, **kw)
    hc = HyperConnectionModule(cfg, 1)
    x = torch.randn(5, 2, 3 * 32)
    h_pre, h_post, h_res = hc.compute_mappings(x)
    assert h_pre.shape == (5, 2, 3) and ((h_pre > 0) & (h_pre < 1)).all() and ((h_post > 0) & (h_post < 2)).all()
    assert torch.allclose(h_res.sum(-1), torch.ones(5, 2, 3), atol=1e-4) and torch.allclose(h_res.sum(-2), torch.ones(5, 2, 3), atol=1e-4)
    # at initialisation the gating factor is small: mappings are nearly input-independent (sigmoid(0) = 1/2, 2 sigmoid(0) = 1, uniform mixing 1/n)
    assert (h_pre - 0.5).abs().max() < 0.05 and (h_post - 1.0).abs().max() < 0.1 and (h_res - 1 / 3).abs().max() < 0.05
    # the pieces compose to x' = H_res x + H_postᵀ F(H_pre x)
    agg, r, p = hc(x)
    assert torch.allclose(agg, (h_pre.unsqueeze(-1) * x.view(5, 2, 3, 32)).sum(2), atol=1e-6)
    y = torch.randn(5, 2, 32)
    out = hc.fused_h_res_h_post_bda(r, x, p, (y, None), 0.0, True).view(5, 2, 3, 32)
    want = torch.einsum("sbij,sbjc->sbic", h_res, x.view(5, 2, 3, 32)) + h_post.unsqueeze(-1) * y.unsqueeze(2)
    assert torch.allclose(out, want, atol=1e-5)
    assert torch.allclose(HyperConnectionModule.output_contract(HyperConnectionModule.input_expand(y, 3), 3), y, atol=1e-6)
    # a GPT model with the n-wide stream: same interface, trains, every mHC parameter receives a gradient
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=16, position_embedding_type="rope")
    names = [n for n, _ in model.named_parameters() if "hyper_connection" in n]
    assert len(names) == 2 * 2 * 5                                                               # 2 layers x 2 sites x (proj, 3 alphas, bias)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    tokens = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16)[None].expand(2, -1)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = model(tokens, pos, None, labels=tokens.roll(-1, 1)).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] and all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in model.named_parameters() if "hyper_connection" in n)
    # validation
    for bad in (dict(mtp_num_layers=1), dict(recompute_granularity="full", recompute_method="uniform", recompute_num_layers=1), dict(pipeline_model_parallel_size=2)):
        try:
            TransformerConfig(enable_mhc_connections=True, **{**kw, **bad})
        except ValueError:
            continue
        raise AssertionError(f"accepted {bad}")
    return True


def test_mhc_module_layer_formula_and_gpt_training():
    assert run_distributed(_mhc_worker, 1) == [True]


def _quant_recipe_worker(rank, world, path):
    """A per-layer precision recipe: fc1 of every layer in FP8 (tensorwise), layer 0 and everything else bf16 — the chosen layers really take the FP8 path
    (their output differs from bf16 by quantisation noise, the others are bit-identical) and the model still trains."""
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.quantization import RecipeConfig, load_quantization_recipe
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    recipe = load_quantization_recipe(path)
    assert isinstance(recipe, RecipeConfig) and recipe.match("decoder.layers.0.mlp.linear_fc1").name == "bf16" and recipe.match("decoder.layers.1.mlp.linear_fc1").name == "fp8"
    kw = dict(num_layers=2, hidden_size=64, num_attention_heads=4, ffn_hidden_size=128, add_bias_linear=False, normalization="RMSNorm", use_cpu_initialization=True,
              hidden_dropout=0.0, attention_dropout=0.0, bias_dropout_fusion=False)
    outs = {}
    for name, qr in (("plain", None), ("recipe", path)):
        model_parallel_cuda_manual_seed(5)
        torch.manual_seed(5)
        m = GPTModel(TransformerConfig(quant_recipe=qr, **kw), get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=16, position_embedding_type="rope")
        if qr:
            assert m.quantized_layers["decoder.layers.1.mlp.linear_fc1"] == "fp8" and m.quantized_layers["decoder.layers.0.mlp.linear_fc1"] == "bf16"
            assert m.quantized_layers["output_layer"] == "bf16" and m.decoder.layers[1].mlp.linear_fc1.quant_config.enabled
        tokens = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(1))
        pos = torch.arange(16)[None].expand(2, -1)
        x = torch.randn(16, 2, 64, generator=torch.Generator().manual_seed(2))
        outs[name] = (m.decoder.layers[0].mlp.linear_fc1(x)[0].detach(), m.decoder.layers[1].mlp.linear_fc1(x)[0].detach())
        loss = m(tokens, pos, None, labels=tokens.roll(-1, 1)).mean()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert torch.equal(outs["plain"][0], outs["recipe"][0])                                    # layer 0 stays bf16 / fp32: identical
    d = (outs["plain"][1] - outs["recipe"][1]).abs().max().item()
    assert 0 < d < 0.1 * outs["plain"][1].abs().max().item()                                  # layer 1 fc1 went through the FP8 GEMM
    return True


def test_per_layer_quantization_recipe(tmp_path):
    p = tmp_path / "recipe.yaml"
    p.write_text("configs:\n  fp8: {recipe: tensorwise, fp8_format: hybrid}\n  bf16: {recipe: none}\nmatchers:\n  - {pattern: 'decoder.layers.0.*', config: bf16}\n"
                 "  - {pattern: '*.linear_fc1', config: fp8}\n  - {pattern: '*', config: bf16}\n")
    assert run_distributed(_quant_recipe_worker, 1, str(p)) == [True]


def test_mup_config_init_and_optimizer_groups():
    """µP: hidden init std shrinks by 1/sqrt(m), embeddings keep the base std, softmax scale is 1/d_head, logits are multiplied by 1/m, and the optimizer trains
    hidden matrices with lr / m and eps / m while vector-like and embedding parameters keep the base values (reference ``transformer_config.py:2530-2605``,
    ``optimizer/__init__.py:get_mup_config_overrides``)."""
    import os

    import torch

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.optimizer import OptimizerConfig, get_megatron_optimizer
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    cfg = TransformerConfig(num_layers=2, hidden_size=256, num_attention_heads=4, use_mup=True, mup_base_hidden_size=64, init_method_std=0.02, use_cpu_initialization=True)
    assert cfg.mup_width_mult == 4.0 and cfg.mup_output_mult == 0.25 and cfg.softmax_scale == 1.0 / 64
    t = torch.empty(512, 512)
    assert abs(float(cfg.init_method(t).std()) - 0.01) < 1e-3 and abs(float(cfg.embedding_init_method(t).std()) - 0.02) < 2e-3
    assert abs(float(cfg.output_layer_init_method(t).std()) - 0.02 / (2.0 * 2.0)) < 1e-3
    base = TransformerConfig(num_layers=2, hidden_size=256, num_attention_heads=4, use_mup=True, use_cpu_initialization=True)
    assert base.mup_width_mult == 1.0 and base.mup_output_mult == 1.0
    own = not torch.distributed.is_initialized()
    if own:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29762")
        torch.distributed.init_process_group("gloo", rank=0, world_size=1)
    if not ps.is_initialized():
        ps.initialize_model_parallel()
    try:
        m = GPTModel(cfg, get_gpt_layer_local_spec(), vocab_size=64, max_sequence_length=16, share_embeddings_and_output_weights=False)
        opt = get_megatron_optimizer(OptimizerConfig(optimizer="adam", lr=1e-2, min_lr=0.0, adam_eps=1e-8, weight_decay=0.0), [m])
        by_param = {id(p): g for g in opt.param_groups for p in g["params"]} if hasattr(opt, "param_groups") else {}
        named = dict(m.named_parameters())
        hidden = by_param[id(named["decoder.layers.0.mlp.linear_fc1.weight"])]
        emb = by_param[id(named["embedding.word_embeddings.weight"])]
        out = by_param[id(named["output_layer.weight"])]
        norm = by_param[id(named["decoder.final_layernorm.weight"])]
        assert hidden["lr_mult"] == 0.25 and abs(hidden["eps"] - 2.5e-9) < 1e-15
        for g in (emb, out, norm):
            assert g["lr_mult"] == 1.0 and g["eps"] == 1e-8
    finally:
        ps.destroy_model_parallel()
        if own:
            torch.distributed.destroy_process_group()
