"""tcgen05/TMA GEMMs (ops/csrc/gemm_sm100.cu) vs torch fp32 reference: 3 layouts x 4 tile/cluster variants."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 256, 64),      # one tile, one k-block
    (256, 512, 256),     # multiple tiles and k-blocks (pipeline wrap-around)
    (1024, 768, 4096),   # QKV @ tp=8
    (1000, 520, 200),    # ragged M / N / K tails
    (8192, 4096, 512),   # proj @ tp=8
    (384, 16032, 1024),  # LM-head shard, N not a multiple of 256
]
VARIANTS = [1, 2, 3, 4, 5, 6]  # 1cta-128x256, 1cta-128x128, 2cta-256x256, 2cta-256x128, 5/6 = 3/4 with the TMA-store (reduce-add when accumulating) epilogue


def _ops(variant):
    from megatron_b200 import ops

    assert ops.has_ext(), f"native extension missing: {ops._EXT_ERR!r}"
    ops.set_gemm_backend(f"tcgen05:{variant}")
    return ops


def _check(out, ref, K):
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 2e-2 * scale + 1e-2, f"max abs err {err} (ref scale {scale}, K={K})"


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_nt(M, N, K, variant):
    ops = _ops(variant)
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    _check(ops.gemm_nt(a, b), a.float() @ b.float().t(), K)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_nn(M, N, K, variant):
    ops = _ops(variant)
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(K, N, device="cuda").bfloat16()
    _check(ops.gemm_nn(a, b), a.float() @ b.float(), K)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("acc", [False, True])
def test_gemm_tn_wgrad(M, N, K, acc, variant):
    ops = _ops(variant)
    torch.manual_seed(0)
    a = torch.randn(K, M, device="cuda").bfloat16()  # dY  [tokens, out]
    b = torch.randn(K, N, device="cuda").bfloat16()  # X   [tokens, in]
    ref = a.float().t() @ b.float()
    if acc:
        main_grad = torch.randn(M, N, device="cuda", dtype=torch.float32)
        ref = ref + main_grad
        out = ops.gemm_tn(a, b, out=main_grad, accumulate=True)
    else:
        out = ops.gemm_tn(a, b)
    _check(out, ref, K)


def test_autotuner_picks_a_candidate_and_is_correct():
    from megatron_b200 import ops
    from megatron_b200.ops import gemm as g

    ops.set_gemm_backend("auto")
    torch.manual_seed(0)
    a = torch.randn(2048, 4096, device="cuda").bfloat16()
    b = torch.randn(3584, 4096, device="cuda").bfloat16()
    ref = a.float() @ b.float().t()
    ours = ops.gemm_nt(a, b).float()
    lib = (a @ b.t()).float()
    assert (ours - ref).abs().max() <= 2 * (lib - ref).abs().max() + 1e-3
    assert g.tuning_report(), "autotuner did not run"
    print(g.tuning_report()[-1])


@pytest.mark.parametrize("M,N,K", [(512, 512, 256), (1000, 768, 1024), (4096, 4096, 4096)])
def test_fp8_gemm_matches_emulation(M, N, K):
    """tcgen05 kind::f8f6f4 GEMM == the same FP8-quantised operands multiplied in fp32."""
    from megatron_b200 import ops
    from megatron_b200.core.fp8_utils import E4M3, E5M2, quantize

    assert hasattr(ops.ext(), "gemm_fp8_nt"), "fp8 GEMM not built"
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    for da, db in ((E4M3, E4M3), (E5M2, E4M3)):
        aq, ai = quantize(a, da)
        bq, bi = quantize(b, db)
        out = ops.ext().gemm_fp8_nt(aq, bq, 1.0, (ai * bi).float())
        ref = (aq.float() @ bq.float().t()) * (ai * bi)
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-2, f"{da}x{db}: rel err {err}"


def test_gemm_is_batch_invariant_in_the_mode():
    """Batch-invariant mode pins ONE tcgen05 variant (no autotuner, no library, no split-K): a row's result does not depend on how many other rows are in the
    batch — bitwise — for the forward (NT) and the dgrad (NN) GEMMs of a linear layer."""
    from megatron_b200 import ops
    from megatron_b200.core.transformer.custom_layers.batch_invariant_kernels import set_batch_invariant_mode

    torch.manual_seed(0)
    K, N = 4096, 1536
    x = torch.randn(4096, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    gy = torch.randn(4096, N, device="cuda").bfloat16()
    with set_batch_invariant_mode(True):
        full_f, full_b = ops.gemm_nt(x, w), ops.gemm_nn(gy, w)
        for m in (1, 8, 129, 1000, 2048):
            assert torch.equal(ops.gemm_nt(x[:m].contiguous(), w), full_f[:m]), f"forward rows differ at batch {m}"
            assert torch.equal(ops.gemm_nn(gy[:m].contiguous(), w), full_b[:m]), f"dgrad rows differ at batch {m}"
        assert torch.equal(ops.gemm_nt(x[1000:1016].contiguous(), w), full_f[1000:1016])      # and not on where in the batch the row sits
    _check(full_f, x.float() @ w.float().t(), K)
