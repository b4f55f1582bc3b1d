"""Numerics of every sm_100a kernel against the plain-PyTorch fp32 reference of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from megatron_b200 import ops

    assert ops.has_ext(), f"native extension missing: {ops._EXT_ERR!r}"
    return ops


def _close(a, b, atol, rtol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), f"{what}: max abs err {err:.4e}"


@pytest.mark.parametrize("rows,H", [(512, 4096), (77, 768), (33, 8192), (16, 256)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_rmsnorm(rows, H, dtype):
    ops = _ops()
    torch.manual_seed(0)
    x = torch.randn(rows, H, device="cuda", dtype=dtype, requires_grad=True)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_()
    g = torch.randn(rows, H, device="cuda", dtype=dtype)
    y = ops.rms_norm(x, w, 1e-5)
    y.backward(g)
    xr, wr = x.detach().float().requires_grad_(), w.detach().float().requires_grad_()
    yr, _ = ops.ref.rms_norm_fwd(xr, wr, 1e-5)
    yr.backward(g.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    _close(y, yr, tol, tol, "y")
    _close(x.grad, xr.grad, tol, tol, "dx")
    _close(w.grad, wr.grad, tol * math.sqrt(rows), tol, "dw")


@pytest.mark.parametrize("rows,H", [(256, 4096), (50, 768)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_layernorm(rows, H, dtype):
    ops = _ops()
    torch.manual_seed(0)
    x = torch.randn(rows, H, device="cuda", dtype=dtype, requires_grad=True)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_()
    b = (0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_()
    g = torch.randn(rows, H, device="cuda", dtype=dtype)
    y = ops.layer_norm(x, w, b, 1e-5)
    y.backward(g)
    xr, wr, br = (t.detach().float().requires_grad_() for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (H,), wr, br, 1e-5)
    yr.backward(g.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    _close(y, yr, tol, tol, "y")
    _close(x.grad, xr.grad, tol, tol, "dx")
    _close(w.grad, wr.grad, tol * math.sqrt(rows), tol, "dw")
    _close(b.grad, br.grad, tol * math.sqrt(rows), tol, "db")


@pytest.mark.parametrize("rows,F", [(1024, 14336), (100, 512)])
@pytest.mark.parametrize("with_bias,with_probs", [(False, False), (True, False), (False, True)])
def test_swiglu(rows, F, with_bias, with_probs):
    ops = _ops()
    torch.manual_seed(0)
    y = torch.randn(rows, 2 * F, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    bias = (0.1 * torch.randn(2 * F, device="cuda")).bfloat16().requires_grad_() if with_bias else None
    probs = torch.rand(rows, 1, device="cuda", dtype=torch.float32, requires_grad=True) if with_probs else None
    g = torch.randn(rows, F, device="cuda", dtype=torch.bfloat16)
    out = ops.swiglu(y, bias, probs)
    out.backward(g)
    yr = y.detach().float().requires_grad_()
    br = bias.detach().float().requires_grad_() if with_bias else None
    pr = probs.detach().clone().requires_grad_() if with_probs else None
    yy = yr + br if with_bias else yr
    a, b = yy.chunk(2, -1)
    ref = torch.nn.functional.silu(a) * b
    if with_probs:
        ref = ref * pr
    ref.backward(g.float())
    _close(out, ref, 3e-2, 2e-2, "out")
    _close(y.grad, yr.grad, 3e-2, 2e-2, "dy")
    if with_bias:
        _close(bias.grad, br.grad, 0.5, 5e-2, "dbias")
    if with_probs:
        _close(probs.grad, pr.grad, 0.5, 3e-2, "dprobs")


@pytest.mark.parametrize("S,B,Hh,D,Drot", [(512, 2, 8, 128, 128), (64, 1, 4, 64, 32)])
def test_rope(S, B, Hh, D, Drot):
    ops = _ops()
    torch.manual_seed(0)
    t = torch.randn(S, B, Hh, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    inv = 1.0 / (10000 ** (torch.arange(0, Drot, 2, device="cuda").float() / Drot))
    fr = torch.outer(torch.arange(S, device="cuda").float(), inv)
    freqs = torch.cat([fr, fr], -1)[:, None, None, :]
    g = torch.randn_like(t)
    out = ops.apply_rope(t, freqs)
    out.backward(g)
    tr = t.detach().float().requires_grad_()
    ref = ops.ref.rope_fwd(tr, freqs)
    ref.backward(g.float())
    _close(out, ref, 2e-2, 2e-2, "out")
    _close(t.grad, tr.grad, 2e-2, 2e-2, "dt")


@pytest.mark.parametrize("rows,V", [(256, 16032), (64, 128256), (10, 1000)])
def test_cross_entropy(rows, V):
    ops = _ops()
    torch.manual_seed(0)
    logits = (3 * torch.randn(rows, V, device="cuda")).bfloat16()
    target = torch.randint(0, V, (rows,), device="cuda")
    lr = logits.detach().float().requires_grad_()
    ref = torch.nn.functional.cross_entropy(lr, target, reduction="none")
    gl = torch.rand(rows, device="cuda")
    ref.backward(gl)
    lg = logits.clone().requires_grad_()
    loss = ops.vocab_parallel_cross_entropy(lg, target)
    _close(loss, ref, 2e-3, 2e-3, "loss")
    loss.backward(gl)
    _close(lg.grad, lr.grad, 2e-3, 2e-2, "dlogits")


def test_cross_entropy_vocab_shards_compose():
    """Two vocab shards combined by hand give the same loss as the full vocab."""
    ops = _ops()
    torch.manual_seed(1)
    rows, V = 32, 4096
    logits = torch.randn(rows, V, device="cuda").bfloat16()
    target = torch.randint(0, V, (rows,), device="cuda")
    ref = torch.nn.functional.cross_entropy(logits.float(), target, reduction="none")
    st = [ops.ext().ce_stats(logits[:, i * V // 2 : (i + 1) * V // 2].contiguous(), target, i * V // 2) for i in range(2)]
    gmax = torch.maximum(st[0][0], st[1][0])
    se = st[0][1] * torch.exp(st[0][0] - gmax) + st[1][1] * torch.exp(st[1][0] - gmax)
    loss = gmax + torch.log(se) - (st[0][2] + st[1][2])
    _close(loss, ref, 2e-3, 2e-3)


def test_multi_tensor_l2norm_and_scale():
    ops = _ops()
    torch.manual_seed(0)
    ts = [torch.randn(n, device="cuda", dtype=dt) for n, dt in [(1000003, torch.float32), (17, torch.bfloat16), (8192 * 5, torch.bfloat16), (1, torch.float32)]]
    n = ops.multi_tensor_l2norm(ts)
    ref = torch.sqrt(sum((t.float() ** 2).sum() for t in ts))
    _close(n, ref, 1e-2, 1e-4)
    before = [t.clone() for t in ts]
    ops.multi_tensor_scale(ts, torch.tensor(0.5, device="cuda"))
    for a, b in zip(ts, before):
        _close(a, b.float() * 0.5, 1e-2, 1e-2)


@pytest.mark.parametrize("gdt", [torch.float32, torch.bfloat16])
def test_fused_adam(gdt):
    ops = _ops()
    torch.manual_seed(0)
    sizes = [4096 * 33 + 5, 7, 8192, 100000]
    p = [torch.randn(n, device="cuda") for n in sizes]
    g = [torch.randn(n, device="cuda").to(gdt) for n in sizes]
    m = [torch.zeros(n, device="cuda") for n in sizes]
    v = [torch.zeros(n, device="cuda") for n in sizes]
    lp = [torch.empty(n, device="cuda", dtype=torch.bfloat16) for n in sizes]
    pr, mr, vr = [t.clone() for t in p], [t.clone() for t in m], [t.clone() for t in v]
    gs = torch.tensor([0.5], device="cuda")
    for step in (1, 2, 3):
        ops.fused_adam(p, g, m, v, lp, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=step, adamw=True, grad_scale=gs)
        for i in range(len(sizes)):
            ops.ref.adam_step(pr[i], g[i], mr[i], vr[i], 1e-2, 0.9, 0.95, 1e-8, 0.1, step, True, 0.5, None)
    for i in range(len(sizes)):
        _close(p[i], pr[i], 1e-5, 1e-5, f"p{i}")
        _close(m[i], mr[i], 1e-6, 1e-5, f"m{i}")
        _close(v[i], vr[i], 1e-6, 1e-5, f"v{i}")
        _close(lp[i], pr[i], 2e-2, 1e-2, f"lowp{i}")
