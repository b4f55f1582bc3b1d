"""Grouped (per-expert) tcgen05 GEMM vs per-expert fp32 matmuls, ragged segments incl. empty experts."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tpe,N,K", [([256, 128, 512, 128], 512, 256), ([100, 0, 333, 77, 1, 260], 384, 320), ([1000, 24], 128, 64)])
def test_grouped_gemm_modes(tpe, N, K):
    from megatron_b200 import ops
    from megatron_b200.ops import grouped

    assert hasattr(ops.ext(), "grouped_gemm_bf16"), "native grouped GEMM not built"
    torch.manual_seed(0)
    E, T = len(tpe), sum(tpe)
    x = torch.randn(T, K, device="cuda").bfloat16()
    w = (0.05 * torch.randn(E, N, K, device="cuda")).bfloat16()
    gy = torch.randn(T, N, device="cuda").bfloat16()
    off = [0]
    for t in tpe:
        off.append(off[-1] + t)
    out = grouped.grouped_gemm_nt(x, w, tpe)
    gx = grouped.grouped_gemm_nn(gy, w, tpe)
    gw = grouped.grouped_gemm_tn(gy, x, tpe, w)
    for e in range(E):
        sl = slice(off[e], off[e + 1])
        ref = x[sl].float() @ w[e].float().t()
        assert torch.allclose(out[sl].float(), ref, atol=0.06, rtol=0.03), f"fwd expert {e}"
        ref = gy[sl].float() @ w[e].float()
        assert torch.allclose(gx[sl].float(), ref, atol=0.3, rtol=0.03), f"dgrad expert {e}"
        ref = gy[sl].float().t() @ x[sl].float()
        assert torch.allclose(gw[e].float(), ref, atol=0.5, rtol=0.03), f"wgrad expert {e} max err {(gw[e].float() - ref).abs().max()}"


def test_moe_row_kernels_match_reference():
    from megatron_b200 import ops

    torch.manual_seed(0)
    T, K, H = 300, 2, 512
    x = torch.randn(T, H, device="cuda").bfloat16().requires_grad_(True)
    idx = torch.randint(0, T, (T * K,), device="cuda")
    scale = torch.rand(T * K, device="cuda").requires_grad_(True)
    out = ops.moe_gather_rows(x, idx, scale)
    ref = x.detach().float()[idx] * scale.detach()[:, None]
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=2e-2)
    out.float().sum().backward()
    gref = torch.zeros(T, H, device="cuda").index_add_(0, idx, scale.detach()[:, None].expand(-1, H).contiguous())
    assert torch.allclose(x.grad.float(), gref, atol=5e-2, rtol=5e-2)
    # combine: every token sums its K expert outputs with weights; one slot dropped (-1)
    y = torch.randn(T * K, H, device="cuda").bfloat16().requires_grad_(True)
    pos = torch.randperm(T * K, device="cuda").view(T, K)
    pos[5, 1] = -1
    w = torch.rand(T, K, device="cuda").requires_grad_(True)
    out = ops.moe_combine_rows(y, pos, w)
    m = (pos >= 0).float() * w.detach()
    ref = (y.detach().float()[pos.clamp(min=0)] * m[..., None]).sum(1)
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=2e-2)
    out.float().sum().backward()
    assert y.grad is not None and w.grad is not None and float(y.grad[pos[5, 0]].abs().sum()) > 0


@pytest.mark.parametrize("E,topk,fn", [(8, 2, "softmax"), (64, 6, "sigmoid"), (256, 8, "softmax")])
def test_fused_topk_router_matches_torch(E, topk, fn):
    from megatron_b200 import ops

    torch.manual_seed(1)
    logits = torch.randn(1000, E, device="cuda")
    bias = torch.randn(E, device="cuda") * 0.1 if fn == "sigmoid" else None
    ids, rmap, tpe, probs = ops.moe_topk_router(logits, topk, fn, pre_softmax=False, expert_bias=bias)
    scores = torch.sigmoid(logits) if fn == "sigmoid" else logits
    key = scores + bias if bias is not None else scores
    rid = torch.topk(key, topk, dim=1).indices
    assert torch.equal(ids.sort(1).values, rid.sort(1).values)
    vals = scores.gather(1, ids)
    rp = torch.softmax(vals, -1) if fn == "softmax" else vals / vals.sum(-1, keepdim=True)
    assert torch.allclose(probs, rp, atol=1e-5)
    assert int(tpe.sum()) == 1000 * topk and torch.equal(rmap.sum(0).int(), tpe)
