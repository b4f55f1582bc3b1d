"""misc_kernels.cu: paged stash copy / pop, batched speculative verify, fused bias-dropout-add vs PyTorch formulations."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_paged_stash_roundtrip_uses_the_kernels():
    from megatron_b200 import ops
    from megatron_b200.core.transformer.moe.paged_stash import PagedStashBuffer, PagedTensor

    torch.manual_seed(0)
    T, H, ps = 1000, 256, 64
    buf = PagedStashBuffer(4096, H, ps, "cuda", torch.bfloat16)
    handles, originals = [], []
    ops.reset_launch_count()
    for n_valid in (1000, 130, 64, 1):
        x = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
        h = PagedTensor(x, torch.tensor(n_valid, device="cuda"))
        h.offload_to_stash(buf)
        handles.append(h)
        originals.append((x, n_valid))
    assert ops.launch_count() == 4
    used = sum(-(-n // ps) for _, n in originals)
    assert buf.free_pages() == buf.num_pages - used
    for h, (x, n) in reversed(list(zip(handles, originals))):
        y = h.reload_from_stash(buf)
        assert torch.equal(y[:n], x[:n]) and not y[n:].any()
    assert buf.free_pages() == buf.num_pages and int(buf.overflow) == 0
    assert ops.launch_count() == 8


@pytest.mark.parametrize("B,k,V,with_draft_probs", [(16, 4, 32000, True), (7, 1, 1000, True), (5, 3, 517, False), (3, 0, 64, False)])
def test_spec_verify_matches_reference(B, k, V, with_draft_probs):
    from megatron_b200.core.inference.speculative import verify_draft_tokens_batched

    torch.manual_seed(1)
    tp = torch.softmax(torch.randn(B, k + 1, V, device="cuda") * 2, -1)
    dp = torch.softmax(torch.randn(B, k, V, device="cuda") * 2 + 0.5 * torch.log(tp[:, :k]), -1) if with_draft_probs else None
    if k:
        dt = torch.multinomial(dp.view(-1, V), 1).view(B, k) if with_draft_probs else torch.where(torch.rand(B, k, device="cuda") < 0.6, tp[:, :k].argmax(-1), torch.randint(0, V, (B, k), device="cuda"))
    else:
        dt = torch.zeros(B, 0, dtype=torch.long, device="cuda")
    ua, us = torch.rand(B, k, device="cuda"), torch.rand(B, device="cuda")
    if not with_draft_probs and k:
        tp = torch.nn.functional.one_hot(tp.argmax(-1), V).float() * 0.7 + tp * 0.3       # peaky target so that some greedy drafts are accepted
    n, nxt = verify_draft_tokens_batched(dt, dp, tp, ua, us)
    n_ref, nxt_ref = verify_draft_tokens_batched(dt.cpu(), None if dp is None else dp.cpu(), tp.cpu(), ua.cpu(), us.cpu())
    assert torch.equal(n.cpu(), n_ref)
    same = nxt.cpu() == nxt_ref
    if not bool(same.all()):                                                                 # fp32 prefix sums in a different order: allow a neighbour at a CDF boundary
        rows = (~same).nonzero().flatten()
        for r in rows.tolist():
            dist = tp[r, n_ref[r]].cpu() if n_ref[r] == k else (tp[r, n_ref[r]].cpu() - (dp[r, n_ref[r]].cpu() if dp is not None else torch.nn.functional.one_hot(dt[r, n_ref[r]].cpu(), V).float())).clamp(min=0)
            cdf = dist.cumsum(0) / dist.sum()
            a, b = sorted((int(nxt[r]), int(nxt_ref[r])))
            assert abs(float(cdf[a]) - float(us[r])) < 1e-4 or abs(float(cdf[b - 1]) - float(us[r])) < 1e-4, (r, a, b)
    if k:
        assert 0 < n.float().mean().item() < k                                                # the case exercises both outcomes


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("with_bias", [True, False])
def test_bias_dropout_add_fused(dtype, with_bias):
    from megatron_b200 import ops

    torch.manual_seed(2)
    s, b, h, p = 512, 4, 1024, 0.1
    x = torch.randn(s, b, h, device="cuda", dtype=dtype, requires_grad=True)
    res = torch.randn(s, b, h, device="cuda", dtype=dtype, requires_grad=True)
    bias = torch.randn(h, device="cuda", dtype=dtype, requires_grad=True) if with_bias else None
    # p = 0 (eval): exact
    y0 = ops.bias_dropout_add(x, bias, res, p, training=False)
    want0 = res.float() + x.float() + (bias.float() if with_bias else 0)
    assert (y0.float() - want0).abs().max().item() <= (2e-2 if dtype == torch.bfloat16 else 1e-6) * want0.abs().max().item()
    # training: the backward regenerates the keep mask from (seed, offset) — read it with a gradient of ones, then check the forward against it
    state = torch.cuda.get_rng_state()
    ops.reset_launch_count()
    y = ops.bias_dropout_add(x, bias, res, p, training=True)
    assert ops.launch_count() == 1
    (m,) = torch.autograd.grad(y, x, torch.ones_like(y), retain_graph=True)
    m = m.float()
    keep_scale = torch.tensor(1 / (1 - p), dtype=dtype).float().item()
    assert bool(((m == 0) | ((m - keep_scale).abs() < 1e-6)).all())
    rate = (m != 0).float().mean().item()
    assert abs(rate - (1 - p)) < 5e-3, rate
    inner = (x.float() + (bias.float() if with_bias else 0))
    want = res.float() + inner * (m != 0) / (1 - p)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert (y.float() - want).abs().max().item() <= tol * want.abs().max().item()
    g = torch.randn_like(y)
    y.backward(g)
    gx_want = g.float() * (m != 0) / (1 - p)
    assert (x.grad.float() - gx_want).abs().max().item() <= tol * gx_want.abs().max().item()
    assert torch.equal(res.grad, g)
    if with_bias:
        assert (bias.grad.float() - x.grad.float().sum((0, 1))).abs().max().item() <= tol * s * b ** 0.5
    # same generator state -> same mask; advanced state -> different mask
    y_next = ops.bias_dropout_add(x.detach(), None if bias is None else bias.detach(), res.detach(), p, training=True)
    assert not torch.equal(y_next, y.detach())
    torch.cuda.set_rng_state(state)
    y_again = ops.bias_dropout_add(x.detach(), None if bias is None else bias.detach(), res.detach(), p, training=True)
    assert torch.equal(y_again, y.detach())
