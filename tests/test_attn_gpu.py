"""tcgen05 flash-attention forward vs an fp32 PyTorch reference (values, log-sum-exp, gradients through the library backward)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, causal, scale):
    sq, b, hq, d = q.shape
    sk, hk = k.shape[0], k.shape[2]
    rep = hq // hk
    qf = q.permute(1, 2, 0, 3).float()
    kf = k.permute(1, 2, 0, 3).float().repeat_interleave(rep, dim=1)
    vf = v.permute(1, 2, 0, 3).float().repeat_interleave(rep, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(torch.ones(sq, sk, dtype=torch.bool, device=q.device).triu(diagonal=1 + sk - sq), float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    o = torch.matmul(torch.softmax(s, dim=-1), vf)
    return o.permute(2, 0, 1, 3), lse


@pytest.mark.parametrize(
    "sq,sk,b,hq,hk,d,causal",
    [
        (256, 256, 1, 2, 2, 128, True),
        (512, 512, 2, 4, 2, 128, True),
        (384, 384, 1, 4, 1, 64, True),      # sq not a multiple of 256: second tile partly/fully out of range
        (200, 200, 1, 2, 2, 128, True),     # ragged: sk not a multiple of 64
        (256, 1000, 1, 2, 1, 128, False),   # cross-attention shaped, key padding in the last block
        (128, 640, 1, 8, 2, 128, True),     # bottom-right aligned causal (chunked prefill)
        (2048, 2048, 1, 8, 2, 128, True),
    ],
)
@pytest.mark.parametrize("variant", [0, 1])
def test_flash_fwd_matches_reference(sq, sk, b, hq, hk, d, causal, variant):
    from megatron_b200 import ops
    _fwd_case(ops, sq, sk, b, hq, hk, d, causal, variant)


@pytest.mark.parametrize("sq,hq,hk", [(2048, 4, 1), (1024, 2, 2), (1152, 2, 1), (1000, 4, 2)])
def test_flash_fwd_mirrored_pairing(sq, hq, hk, monkeypatch):
    """Few-head grids (TP=8): a CTA owns query tiles (x, T-1-x); even / odd tile counts and a ragged last tile."""
    from megatron_b200 import ops

    monkeypatch.setenv("MB200_FA_PAIR_MODE", "1")
    _fwd_case(ops, sq, sq, 1, hq, hk, 128, True, 1)


def _fwd_case(ops, sq, sk, b, hq, hk, d, causal, variant):

    torch.manual_seed(0)
    q = torch.randn(sq, b, hq, d, device="cuda").bfloat16()
    k = torch.randn(sk, b, hk, d, device="cuda").bfloat16()
    v = torch.randn(sk, b, hk, d, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(d)
    o, lse = ops.ext().flash_attn_fwd(q, k, v, causal, scale, variant)   # 0: P via smem, 1: P kept in TMEM (TS MMA)
    ro, rlse = _ref(q, k, v, causal, scale)
    assert torch.isfinite(o.float()).all()
    err = (o.float() - ro).abs().max().item()
    assert err < 2e-2, f"out max err {err}"
    lerr = (lse - rlse).abs().max().item()
    assert lerr < 2e-2, f"lse max err {lerr}"


def test_flash_fwd_consumes_fused_qkv_views_in_place():
    """k / v sliced out of a fused QKV projection ([s, b, g, (r+2)*d]) are read through strided TMA maps, no copies."""
    from megatron_b200 import ops

    torch.manual_seed(1)
    s, b, g, r, d = 512, 2, 2, 4, 128
    mixed = torch.randn(s, b, g, (r + 2) * d, device="cuda").bfloat16()
    qv, k, v = torch.split(mixed, [r * d, d, d], dim=3)
    q = qv.reshape(s, b, g * r, d)
    assert not k.is_contiguous() and not v.is_contiguous()
    o, _ = ops.ext().flash_attn_fwd(q, k, v, True, 0.088)
    ro, _ = _ref(q, k, v, True, 0.088)
    assert (o.float() - ro).abs().max().item() < 2e-2


def test_flash_attention_autograd_matches_reference():
    from megatron_b200 import ops

    torch.manual_seed(2)
    sq, b, hq, hk, d = 512, 1, 8, 2, 128
    q = torch.randn(sq, b, hq, d, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(sq, b, hk, d, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(sq, b, hk, d, device="cuda").bfloat16().requires_grad_(True)
    go = torch.randn(sq, b, hq, d, device="cuda").bfloat16()
    ops.set_attention_impl("native")
    try:
        o = ops.flash_attention(q, k, v, causal=True)
        o.backward(go)
    finally:
        ops.set_attention_impl("auto")
    g_native = [t.grad.float().clone() for t in (q, k, v)]
    for t in (q, k, v):
        t.grad = None
    ro, _ = _ref(q, k, v, True, 1.0 / math.sqrt(d))
    ro.backward(go.float())
    for name, a, t in zip("qkv", g_native, (q, k, v)):
        ref = t.grad.float()
        err = (a - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        assert err < 3e-2, f"d{name} rel err {err}"


@pytest.mark.parametrize("split_heads", [0, 1])
@pytest.mark.parametrize("sq,sk,b,hq,hk,causal", [(512, 512, 1, 8, 2, True), (384, 384, 2, 4, 4, True), (256, 640, 1, 4, 2, False), (128, 640, 1, 4, 1, True), (1000, 1000, 1, 2, 1, True),
                                                  (1024, 1024, 1, 4, 1, True), (2048, 2048, 1, 8, 2, True), (330, 330, 1, 2, 2, False)])
def test_flash_bwd_native_matches_reference(sq, sk, b, hq, hk, causal, split_heads):
    """tcgen05 backward (delta kernel -> dK/dV in TMEM + dS spilled by TMA -> dQ = dS K GEMM kernel; fused-heads and split-heads grids)
    vs fp32 autograd of the reference attention."""
    from megatron_b200 import ops

    assert hasattr(ops.ext(), "flash_attn_bwd"), "native attention backward not built"
    torch.manual_seed(3)
    d = 128
    q = torch.randn(sq, b, hq, d, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(sk, b, hk, d, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(sk, b, hk, d, device="cuda").bfloat16().requires_grad_(True)
    go = torch.randn(sq, b, hq, d, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(d)
    ro, _ = _ref(q, k, v, causal, scale)
    ro.backward(go.float())
    refs = [t.grad.float().clone() for t in (q, k, v)]
    o, lse = ops.ext().flash_attn_fwd(q.detach(), k.detach(), v.detach(), causal, scale, 1)
    # poison the allocator's free blocks so that a read of a never-written scratch tile shows up as NaN
    junk = torch.full((64 << 20,), float("nan"), device="cuda")
    del junk
    dq, dk, dv = ops.ext().flash_attn_bwd(go, q.detach(), k.detach(), v.detach(), o, lse, causal, scale, split_heads)
    torch.cuda.synchronize()
    for name, a, r in zip("qkv", (dq, dk, dv), refs):
        assert torch.isfinite(a.float()).all(), f"d{name} has non-finite values"
        err = (a.float() - r).abs().max().item() / (r.abs().max().item() + 1e-6)
        assert err < 3e-2, f"d{name} rel err {err}"


def test_flash_attention_autograd_native_backward():
    """ops.flash_attention end to end with the native backward selected (the default of this build): strided q/k/v views of a fused QKV tensor."""
    from megatron_b200 import ops

    torch.manual_seed(5)
    s, b, g, r, d = 1024, 1, 2, 4, 128
    mixed = torch.randn(s, b, g, (r + 2) * d, device="cuda").bfloat16().requires_grad_(True)
    q, k, v = torch.split(mixed, [r * d, d, d], dim=3)
    q = q.reshape(s, b, g * r, d)
    scale = 1.0 / math.sqrt(d)
    old = ops._ATTN_BWD_IMPL
    ops._ATTN_BWD_IMPL = "native"
    try:
        out = ops.flash_attention(q, k, v, causal=True, scale=scale)
        go = torch.randn_like(out)
        out.backward(go)
    finally:
        ops._ATTN_BWD_IMPL = old
    got = mixed.grad.float().clone()
    mixed.grad = None
    q2, k2, v2 = torch.split(mixed, [r * d, d, d], dim=3)
    ro, _ = _ref(q2.reshape(s, b, g * r, d), k2, v2, True, scale)
    ro.backward(go.float())
    ref = mixed.grad.float()
    err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
    assert err < 3e-2, err


def _band_ref(q, k, v, scale, window, cu):
    from megatron_b200.ops import reference as ref

    return ref.attention_fwd(q.float(), k.float(), v.float(), True, scale, window, cu)


@pytest.mark.parametrize("split_heads", [0, 1])
@pytest.mark.parametrize(
    "sq,sk,hq,hk,window,cu",
    [
        (1024, 1024, 4, 2, (200, 0), None),                  # sliding window narrower than a tile pair
        (2048, 2048, 4, 1, (511, 0), None),
        (1000, 1000, 2, 2, (64, 0), None),                   # ragged, window of one key block
        (384, 640, 2, 1, (100, 0), None),                    # bottom-right aligned window
        (1024, 1024, 4, 2, None, [0, 300, 301, 640, 1024]),  # packed sequences incl. a length-1 one and boundaries off the tile grid
        (2048, 2048, 2, 1, None, [0, 128, 1024, 2000]),      # trailing pad tokens (cu[-1] < t)
        (1536, 1536, 2, 2, (150, 0), [0, 700, 1536]),        # both
    ],
)
def test_flash_band_masks_native_fwd_bwd(sq, sk, hq, hk, window, cu, split_heads):
    """Sliding window / packed sequences run INSIDE the tcgen05 kernels as a monotone band (row_lo / col_hi): forward, dK/dV kernel and dQ kernel vs
    fp32 autograd of the dense-masked reference."""
    from megatron_b200 import ops

    torch.manual_seed(11)
    d = 128
    q = torch.randn(sq, 1, hq, d, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(sk, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(sk, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    go = torch.randn(sq, 1, hq, d, device="cuda").bfloat16()
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32) if cu is not None else None
    scale = 1.0 / math.sqrt(d)
    ro = _band_ref(q, k, v, scale, window, cu_t)
    ro.backward(go.float())
    refs = [t.grad.float().clone() for t in (q, k, v)]
    lo, hi = ops.attention_band(sq, sk, window, cu_t, q.device)
    o, lse = ops.ext().flash_attn_fwd(q.detach(), k.detach(), v.detach(), True, scale, 1, lo)
    err = (o.float() - ro.float()).abs().max().item()
    assert torch.isfinite(o.float()).all() and err < 2e-2, f"forward abs err {err}"
    junk = torch.full((64 << 20,), float("nan"), device="cuda")
    del junk
    dq, dk, dv = ops.ext().flash_attn_bwd(go, q.detach(), k.detach(), v.detach(), o, lse, True, scale, split_heads, 0, lo, hi)
    torch.cuda.synchronize()
    for name, a, r in zip("qkv", (dq, dk, dv), refs):
        assert torch.isfinite(a.float()).all(), f"d{name} has non-finite values"
        e = (a.float() - r).abs().max().item() / (r.abs().max().item() + 1e-6)
        assert e < 3e-2, f"d{name} rel err {e}"


def test_packed_sequences_through_dot_product_attention():
    """ops.flash_attention(cu_seqlens=...) == running every packed sequence on its own (values and gradients), and the native path was taken."""
    from megatron_b200 import ops

    torch.manual_seed(2)
    lens = [384, 129, 511]
    t, hq, hk, d = sum(lens), 4, 2, 128
    cu = torch.tensor([0, 384, 513, 1024], device="cuda", dtype=torch.int32)
    q, k, v = (torch.randn(t, 1, h, d, device="cuda").bfloat16().requires_grad_(True) for h in (hq, hk, hk))
    ops.reset_launch_count()
    out = ops.flash_attention(q, k, v, causal=True, cu_seqlens=cu)
    assert ops.launch_count() >= 1, "packed attention fell off the native kernels"
    go = torch.randn_like(out)
    out.backward(go)
    got = [out.float()] + [x.grad.float().clone() for x in (q, k, v)]
    for x in (q, k, v):
        x.grad = None
    outs, s0 = [], 0
    for n in lens:
        o = ops.flash_attention(q[s0:s0 + n], k[s0:s0 + n], v[s0:s0 + n], causal=True) if n >= 128 else _band_ref(q[s0:s0 + n], k[s0:s0 + n], v[s0:s0 + n], 1 / math.sqrt(d), None, None).bfloat16()
        outs.append(o)
        s0 += n
    sep = torch.cat(outs, 0)
    sep.backward(go)
    want = [sep.float()] + [x.grad.float() for x in (q, k, v)]
    for name, a, r in zip(["out", "dq", "dk", "dv"], got, want):
        e = (a - r).abs().max().item() / (r.abs().max().item() + 1e-6)
        assert e < 3e-2, f"{name} rel err {e}"


@pytest.mark.parametrize("hq,hk", [(4, 1), (2, 2)])
def test_ring_attention_blocks_on_native_kernels(hq, hk):
    """The context-parallel ring's block primitives on the tcgen05 kernels (one rank = two zig-zag chunks: a full pair, two diagonal pairs, merge by
    log-sum-exp; backward of every pair from the MERGED output / lse) == plain causal attention over the whole local sequence."""
    from megatron_b200 import ops
    from megatron_b200.parallel.context_parallel import _RingAttnFn, _native_block_ok

    torch.manual_seed(8)
    s, d = 1024, 128
    q = torch.randn(s, 1, hq, d, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    go = torch.randn(s, 1, hq, d, device="cuda").bfloat16()
    assert _native_block_ok(q[:512], k[:512])
    scale = 1.0 / math.sqrt(d)
    ops.reset_launch_count()
    out = _RingAttnFn.apply(q, k, v, scale, True, 0, 1, lambda x, reverse: x)
    out.backward(go)
    assert ops.launch_count() >= 3 + 9                       # 3 forward pairs + 3 backward pairs (3 kernels each)
    got = [out.float()] + [t.grad.float().clone() for t in (q, k, v)]
    for t in (q, k, v):
        t.grad = None
    ro, _ = _ref(q, k, v, True, scale)
    ro.backward(go.float())
    want = [ro.float()] + [t.grad.float() for t in (q, k, v)]
    for name, a, r in zip(["out", "dq", "dk", "dv"], got, want):
        e = (a - r).abs().max().item() / (r.abs().max().item() + 1e-6)
        assert e < 3e-2, f"{name} rel err {e}"


def test_gpt_model_packed_documents_on_gpu():
    """A bf16 GPT (head dim 128) on packed documents: cu_seqlens → band mask in the tcgen05 attention kernels + per-document RoPE rows; per-token losses equal
    running every document on its own, and the native attention path was taken."""
    import os

    import torch.distributed as dist
    import torch.nn.functional as F

    from megatron_b200 import ops
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.packed_seq_params import packed_seq_params_from_documents
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29734")
        dist.init_process_group("nccl", rank=0, world_size=1)
    mine = not ps.model_parallel_is_initialized()
    if mine:
        ps.initialize_model_parallel()
    try:
        model_parallel_cuda_manual_seed(9)
        torch.manual_seed(9)
        cfg = TransformerConfig(num_layers=2, hidden_size=256, num_attention_heads=2, ffn_hidden_size=512, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                                normalization="RMSNorm", hidden_dropout=0.0, attention_dropout=0.0, bf16=True, params_dtype=torch.bfloat16)
        model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=512, max_sequence_length=1024, position_embedding_type="rope").cuda()
        eod, b, s = 511, 2, 512
        tokens = torch.randint(0, 510, (b, s), device="cuda")
        for r, c in ((0, 199), (0, 330), (1, 127), (1, 511)):
            tokens[r, c] = eod
        labels = tokens.roll(-1, 1)
        bounds = [(0, 0, 200), (0, 200, 331), (0, 331, 512), (1, 0, 128), (1, 128, 512)]
        pid = torch.cat([torch.cat([torch.arange(hi - lo) for r, lo, hi in bounds if r == i]) for i in range(b)]).view(b, s).cuda()
        want = torch.zeros(b, s, device="cuda")
        with torch.no_grad():
            for r, lo, hi in bounds:
                want[r, lo:hi] = model(tokens[r:r + 1, lo:hi], torch.arange(hi - lo, device="cuda")[None], None, labels=labels[r:r + 1, lo:hi])[0].float()
            psp = packed_seq_params_from_documents(tokens, eod)
            assert psp.cu_seqlens_q.tolist() == [0, 200, 331, 512, 640, 1024]
            ops.reset_launch_count()
            got = model(tokens.reshape(1, -1), pid.reshape(1, -1), None, labels=labels.reshape(1, -1), packed_seq_params=psp).reshape(b, s).float()
            assert ops.launch_count() > 0
        assert (got - want).abs().max().item() < 5e-2 * want.abs().max().item(), (got - want).abs().max().item()
    finally:
        if mine:
            ps.destroy_model_parallel()
