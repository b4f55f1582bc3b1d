"""Gated activations (core/fusions/_gated.py): hand-derived backward == autograd of the plain formula, incl. clamp / offset / weights."""
import pytest
import torch
import torch.nn.functional as F


def _plain(y, bias, w, kind, clamp, offset, alpha):
    x = y if bias is None else y + bias
    a, b = x.chunk(2, -1)
    if clamp is not None:
        a, b = a.clamp(max=clamp), b.clamp(min=-clamp, max=clamp)
    act = {"silu": F.silu, "gelu": lambda t: F.gelu(t, approximate="tanh"), "quick_gelu": lambda t: t * torch.sigmoid(alpha * t)}[kind]
    out = act(a) * (b + offset)
    return out * w if w is not None else out


@pytest.mark.parametrize("kind", ["silu", "gelu", "quick_gelu"])
@pytest.mark.parametrize("clamp,offset", [(None, 0.0), (1.5, 1.0)])
@pytest.mark.parametrize("with_bias,with_w", [(False, False), (True, True)])
def test_gated_activation_matches_autograd(kind, clamp, offset, with_bias, with_w):
    from megatron_b200.core.fusions._gated import gated_activation

    torch.manual_seed(0)
    y = (2 * torch.randn(7, 16, dtype=torch.float64)).requires_grad_()
    bias = torch.randn(16, dtype=torch.float64, requires_grad=True) if with_bias else None
    w = torch.rand(7, 1, dtype=torch.float64, requires_grad=True) if with_w else None
    ins = [t for t in (y, bias, w) if t is not None]
    want = _plain(y, bias, w, kind, clamp, offset, 1.702)
    got = gated_activation(y, bias, w, kind, clamp, offset, 1.702)
    assert torch.allclose(got, want, atol=1e-6)
    g = torch.randn_like(want)
    for a, b in zip(torch.autograd.grad(got, ins, g), torch.autograd.grad(want, ins, g)):
        assert torch.allclose(a, b, atol=1e-5), (kind, clamp, float((a - b).abs().max()))


def test_reference_named_entry_points():
    from megatron_b200.core.fusions import fused_bias_geglu as G, fused_bias_swiglu as S

    y, b, w = torch.randn(2, 3, 8), torch.randn(8), torch.rand(6, 1)
    g = torch.randn(2, 3, 4)
    assert torch.allclose(S.bias_swiglu_impl(y, b, clamp_value=0.5), _plain(y, b, None, "silu", 0.5, 0.0, 1.702), atol=1e-6)
    assert torch.allclose(S.bias_swiglu_impl(y, b), _plain(y, b, None, "silu", None, 0.0, 1.702), atol=1e-6)
    assert torch.allclose(S.weighted_bias_swiglu_impl(y.view(6, 8), None, w, clamp_value=1.0), _plain(y.view(6, 8), None, w, "silu", 1.0, 0.0, 1.702), atol=1e-6)
    assert torch.allclose(G.weighted_bias_quick_geglu_impl(y.view(6, 8), b, w, linear_offset=1.0, clamp_value=7.0), _plain(y.view(6, 8), b, w, "quick_gelu", 7.0, 1.0, 1.702), atol=1e-6)
    yy = y.clone().requires_grad_()
    (gy,) = torch.autograd.grad(_plain(yy, None, None, "silu", None, 0.0, 1.702), yy, g)
    assert torch.allclose(S.swiglu_back(g, y), gy, atol=1e-5)
    (gy,) = torch.autograd.grad(_plain(yy, None, None, "gelu", None, 0.0, 1.702), yy, g)
    assert torch.allclose(G.geglu_back(g, y), gy, atol=1e-5)
    out = S.BiasSwiGLUFunction.apply(yy, b.clone().requires_grad_(), True, False, None)      # fp8 input store: coarse but finite
    out.sum().backward()
    assert torch.isfinite(yy.grad).all()
