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
