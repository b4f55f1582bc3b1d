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
