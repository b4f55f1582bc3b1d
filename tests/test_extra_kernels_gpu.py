"""Numerics of the second kernel batch against plain PyTorch fp32 references (csrc/extra_kernels.cu, residual-fused norm in csrc/norm.cu)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-6
    assert err <= tol * ref, f"max err {err} vs ref max {ref} (tol {tol})"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("rows,H", [(333, 4096), (64, 1024), (17, 8192)])
def test_add_rms_norm_fwd_bwd(dtype, tol, rows, H):
    from megatron_b200 import ops

    torch.manual_seed(0)
    x = torch.randn(rows, H, device="cuda", dtype=dtype, requires_grad=True)
    r = torch.randn(rows, H, device="cuda", dtype=dtype, requires_grad=True)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_()
    gy, gh = torch.randn(rows, H, device="cuda", dtype=dtype), torch.randn(rows, H, device="cuda", dtype=dtype)
    n0 = ops.launch_count()
    y, h = ops.add_rms_norm(x, r, w, 1e-5)
    torch.autograd.backward([y, h], [gy, gh])
    assert ops.launch_count() - n0 == 3
    xf, rf, wf = (t.detach().float().requires_grad_() for t in (x, r, w))
    hf = (xf + rf).to(dtype).float() if dtype != torch.float32 else xf + rf        # the kernel rounds h once
    hf2 = xf + rf
    yf = hf2 * torch.rsqrt(hf2.pow(2).mean(-1, keepdim=True) + 1e-5) * wf
    torch.autograd.backward([yf, hf2], [gy.float(), gh.float()])
    _close(h, hf, tol)
    _close(y, yf, tol)
    _close(x.grad, xf.grad, tol)
    _close(r.grad, rf.grad, tol)
    _close(w.grad, wf.grad, 3 * tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_rope_thd_and_fused_qkv(dtype, tol):
    from megatron_b200 import ops
    from megatron_b200.ops import reference as ref

    torch.manual_seed(1)
    cu = torch.tensor([0, 100, 613, 1024], device="cuda", dtype=torch.int32)
    t = torch.randn(1024, 8, 128, device="cuda", dtype=dtype, requires_grad=True)
    f = torch.randn(1024, 64, device="cuda")
    freqs = torch.cat([f, f], dim=-1)          # Megatron convention: both rotary halves share the angle table
    out = ops.apply_rope_thd(t, cu, freqs)
    g = torch.randn_like(out)
    out.backward(g)
    pos = ops.positions_from_cu_seqlens(cu, 1024).long()
    tf = t.detach().float().requires_grad_()
    of = ref.rope_fwd(tf.unsqueeze(1), freqs[pos][:, None, None, :]).squeeze(1)
    of.backward(g.float())
    _close(out, of, tol)
    _close(t.grad, tf.grad, tol)
    # fused QKV: 4 groups x (2 q + k + v), partial rotary (64 of 128)
    s, b, ng, qpg, d = 257, 2, 4, 2, 128
    qkv = torch.randn(s, b, ng, (qpg + 2) * d, device="cuda", dtype=dtype, requires_grad=True)
    f = torch.randn(s, 32, device="cuda")
    fr = torch.cat([f, f], dim=-1)
    o = ops.apply_rope_qkv(qkv, fr, qpg, d)
    go = torch.randn_like(o)
    o.backward(go)
    qf = qkv.detach().float().requires_grad_()
    x5 = qf.view(s, b, ng, qpg + 2, d)
    rot = ref.rope_fwd(x5[:, :, :, : qpg + 1].reshape(s, b, ng * (qpg + 1), d), fr[:, None, None, :]).view(s, b, ng, qpg + 1, d)
    of = torch.cat([rot, x5[:, :, :, qpg + 1 :]], dim=3).reshape(qkv.shape)
    of.backward(go.float())
    _close(o, of, tol)
    _close(qkv.grad, qf.grad, tol)
    assert torch.equal(o.view(s, b, ng, qpg + 2, d)[:, :, :, -1], qkv.view(s, b, ng, qpg + 2, d)[:, :, :, -1])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("K,with_left,silu", [(4, False, True), (4, True, True), (3, True, False), (2, False, True)])
def test_causal_conv1d_fwd_bwd(dtype, tol, K, with_left, silu):
    from megatron_b200 import ops
    from megatron_b200.ops.extra import _conv1d_ref

    torch.manual_seed(2)
    b, d, l = 2, 96, 777
    x = torch.randn(b, d, l, device="cuda", dtype=dtype, requires_grad=True)
    w = (0.5 * torch.randn(d, K, device="cuda")).to(dtype).requires_grad_()
    bias = (0.1 * torch.randn(d, device="cuda")).to(dtype).requires_grad_()
    left = torch.randn(b, d, K - 1, device="cuda", dtype=dtype, requires_grad=True) if with_left else None
    y = ops.causal_conv1d(x, w, bias, left, silu=silu)
    g = torch.randn_like(y)
    y.backward(g)
    xs = [t.detach().float().requires_grad_() if t is not None else None for t in (x, w, bias, left)]
    yf = _conv1d_ref(xs[0], xs[1], xs[2], xs[3], silu)
    yf.backward(g.float())
    _close(y, yf, tol)
    _close(x.grad, xs[0].grad, tol)
    _close(w.grad, xs[1].grad, 2 * tol)
    _close(bias.grad, xs[2].grad, 2 * tol)
    if with_left:
        _close(left.grad, xs[3].grad, tol)


def test_ssd_state_passing_step_and_chunk_scan():
    from megatron_b200 import ops
    from megatron_b200.core.ssm.ssd import ssd_chunk_scan, ssd_reference, ssd_step
    from megatron_b200.ops.extra import _state_passing_ref

    torch.manual_seed(3)
    b, c, h, p, n = 2, 9, 6, 64, 128
    states = torch.randn(b, c, h, p, n, device="cuda", requires_grad=True)
    decay = (-torch.rand(b, h, c, device="cuda")).requires_grad_()
    init = torch.randn(b, h, p, n, device="cuda", requires_grad=True)
    prev, fin = ops.ssd_state_passing(states, decay, init)
    gp, gf = torch.randn_like(prev), torch.randn_like(fin)
    torch.autograd.backward([prev, fin], [gp, gf])
    s2, d2, i2 = (t.detach().clone().requires_grad_() for t in (states, decay, init))
    prev_r, fin_r = _state_passing_ref(s2, d2, i2)
    torch.autograd.backward([prev_r, fin_r], [gp, gf])
    _close(prev, prev_r, 1e-5)
    _close(fin, fin_r, 1e-5)
    _close(states.grad, s2.grad, 1e-5)
    _close(init.grad, i2.grad, 1e-5)
    _close(decay.grad, d2.grad, 1e-4)
    # decode step kernel vs the einsum definition
    g = 2
    x = torch.randn(b, h, p, device="cuda", dtype=torch.bfloat16)
    dt = torch.rand(b, h, device="cuda") * 0.5
    A = -torch.rand(h, device="cuda")
    B, C = torch.randn(b, g, n, device="cuda", dtype=torch.bfloat16), torch.randn(b, g, n, device="cuda", dtype=torch.bfloat16)
    D = torch.randn(h, device="cuda")
    st = torch.randn(b, h, p, n, device="cuda")
    with torch.enable_grad():                       # einsum path
        y_ref, st_ref = ssd_step(x, dt, A, B, C, st.clone(), D)
    with torch.no_grad():                           # kernel path
        y_k, st_k = ssd_step(x, dt, A, B, C, st.clone(), D)
    _close(y_k, y_ref, 2e-2)
    _close(st_k, st_ref, 1e-5)
    # whole chunked scan on the GPU (state passing kernel inside) vs the sequential definition
    l, hh, pp, nn = 200, 4, 16, 32
    xx = torch.randn(1, l, hh, pp, device="cuda")
    dtt = torch.rand(1, l, hh, device="cuda") * 0.3 + 0.01
    AA = -torch.rand(hh, device="cuda") - 0.1
    BB, CC = torch.randn(1, l, 2, nn, device="cuda"), torch.randn(1, l, 2, nn, device="cuda")
    y1, f1 = ssd_chunk_scan(xx, dtt, AA, BB, CC, chunk_size=64, return_final_states=True)
    y0, f0 = ssd_reference(xx, dtt, AA, BB, CC)
    _close(y1, y0, 1e-3)
    _close(f1, f0, 1e-3)


def test_mxfp8_quantize_matches_reference_and_bounds_error():
    from megatron_b200 import ops

    torch.manual_seed(4)
    x = (torch.randn(257, 4096, device="cuda") * torch.exp(torch.randn(257, 1, device="cuda") * 3)).to(torch.bfloat16)
    x[3, :64] = 0
    q, sf = ops.mxfp8_quantize(x)
    q_ref, sf_ref = ops.mxfp8_quantize_reference(x)
    assert torch.equal(sf, sf_ref)
    assert (q != q_ref).float().mean().item() < 1e-3           # round-to-nearest ties may differ between cvt.rn.satfinite and the torch cast
    d = ops.mxfp8_dequantize(q, sf).float()
    blk = x.float().view(257, -1, 32)
    err = (d.view(257, -1, 32) - blk).abs().amax(-1)
    assert (err <= blk.abs().amax(-1) * 2 ** -3 + 1e-30).all()  # e4m3: 3 mantissa bits, block max lands in [256, 448)
    assert torch.all(d[3, :64] == 0)


@pytest.mark.parametrize("tile", [128, 256])
@pytest.mark.parametrize("M,N,K", [(256, 256, 512), (300, 392, 1024), (1024, 4096, 4096)])
def test_mxfp8_block_scaled_gemm(M, N, K, tile):
    from megatron_b200 import ops

    torch.manual_seed(5)
    # per-block dynamic range: the block scales really matter (a per-tensor scale would flush the small blocks to zero)
    a = (torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-6, 7, (M, K // 32), device="cuda").float()).repeat_interleave(32, dim=1)).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-6, 7, (N, K // 32), device="cuda").float()).repeat_interleave(32, dim=1)).to(torch.bfloat16)
    aq, asf = ops.mxfp8_quantize(a)
    bq, bsf = ops.mxfp8_quantize(b)
    n0 = ops.launch_count()
    c = ops.gemm_mxfp8_nt(aq, asf, bq, bsf, tile=tile)
    assert ops.launch_count() == n0 + 1 and c.shape == (M, N) and c.dtype == torch.bfloat16
    ref = ops.mxfp8_dequantize(aq, asf).float() @ ops.mxfp8_dequantize(bq, bsf).float().t()
    _close(c, ref, 1e-2)
    # and the quantised product tracks the bf16 product to fp8 accuracy
    full = a.float() @ b.float().t()
    assert ((c.float() - full).norm() / full.norm()).item() < 0.06


@pytest.mark.parametrize("M,N,K", [(256, 256, 512), (300, 392, 1024), (1024, 2048, 4096)])
def test_nvfp4_block_scaled_gemm(M, N, K):
    from megatron_b200 import ops
    from megatron_b200.core.fp4_utils import dequantize_nvfp4, quantize_nvfp4

    torch.manual_seed(6)
    a = torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-3, 4, (M, K // 16), device="cuda").float()).repeat_interleave(16, dim=1)
    b = torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-3, 4, (N, K // 16), device="cuda").float()).repeat_interleave(16, dim=1)
    qa, qb = quantize_nvfp4(a), quantize_nvfp4(b)
    n0 = ops.launch_count()
    c = ops.gemm_nvfp4_nt(*qa, *qb)
    assert ops.launch_count() == n0 + 1 and c.shape == (M, N) and c.dtype == torch.bfloat16
    ref = dequantize_nvfp4(*qa, torch.float32) @ dequantize_nvfp4(*qb, torch.float32).t()
    _close(c, ref, 1e-2)
    full = a @ b.t()
    assert ((c.float() - full).norm() / full.norm()).item() < 0.2            # 4-bit payload: ~10 % relative error per operand


def test_nvfp4_quantise_kernel_and_w4a4_linear():
    from megatron_b200 import ops
    from megatron_b200.core.fp4_utils import dequantize_nvfp4, nvfp4_linear, quantize_nvfp4

    torch.manual_seed(7)
    x = (torch.randn(300, 1024, device="cuda") * torch.exp2(torch.randint(-4, 5, (300, 64), device="cuda").float()).repeat_interleave(16, dim=1)).to(torch.bfloat16)
    x[5, :32] = 0
    q, sf, t = ops.nvfp4_quantize(x)
    codes_ref, sf_ref, t_ref = quantize_nvfp4(x)
    assert torch.allclose(t, t_ref.reshape(1)) and q.shape == (300, 512)
    assert (sf.view(torch.uint8) != sf_ref.view(torch.uint8)).float().mean().item() < 2e-3            # e4m3 rounding of the block scale
    codes = ops.nvfp4_unpack(q)
    same_scale = (sf.view(torch.uint8) == sf_ref.view(torch.uint8)).repeat_interleave(16, dim=1)
    assert ((codes != codes_ref) & same_scale).float().mean().item() < 0.02                           # ties round to even on the hardware, to the lower code in the reference
    d = dequantize_nvfp4(codes, sf, t, torch.float32)
    d_ref = dequantize_nvfp4(codes_ref, sf_ref, t_ref, torch.float32)
    xf = x.float()
    assert (d - xf).norm() <= 1.02 * (d_ref - xf).norm() and torch.all(d[5, :32] == 0)
    w = torch.randn(512, 1024, device="cuda") * 0.05
    y = nvfp4_linear(x.view(3, 100, 1024), quantize_nvfp4(w))
    ref = xf @ w.t()
    assert y.shape == (3, 100, 512) and ((y.float().view(300, 512) - ref).norm() / ref.norm()).item() < 0.2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_scaled_masked_softmax_and_activations(dtype, tol):
    from megatron_b200 import ops
    from megatron_b200.ops import reference as ref

    torch.manual_seed(8)
    b, h, sq, sk = 2, 3, 70, 133
    x = torch.randn(b, h, sq, sk, device="cuda", dtype=dtype, requires_grad=True)
    mask = torch.rand(b, 1, sq, sk, device="cuda") > 0.8
    for m, causal in ((mask, False), (None, True), (None, False)):
        y = ops.scaled_masked_softmax(x, m, 0.37, causal=causal)
        g = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, g)
        xf = x.detach().float().requires_grad_()
        yf = ref.scaled_masked_softmax(xf, m, 0.37, causal=causal)
        (gxf,) = torch.autograd.grad(yf, xf, g.float())
        _close(y, yf, tol)
        _close(gx, gxf, tol)
        assert torch.allclose(y.float().sum(-1), torch.ones(b, h, sq, device="cuda"), atol=2e-2 if dtype == torch.bfloat16 else 1e-5)
    z = torch.randn(37, 4, 256, device="cuda", dtype=dtype, requires_grad=True)
    for fn, rf in ((ops.squared_relu, lambda t: torch.relu(t) ** 2), (ops.quick_geglu, lambda t: (lambda a, bb: a * torch.sigmoid(1.702 * a) * bb)(*t.chunk(2, -1)))):
        out = fn(z)
        g = torch.randn_like(out)
        (gz,) = torch.autograd.grad(out, z, g)
        zf = z.detach().float().requires_grad_()
        of = rf(zf)
        (gzf,) = torch.autograd.grad(of, zf, g.float())
        _close(out, of, tol)
        _close(gz, gzf, tol)
