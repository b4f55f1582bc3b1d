"""Band masks (sliding window / packed THD sequences): the (row_lo, col_hi) arrays the native kernels consume describe exactly the dense mask; packed
sequences through DotProductAttention equal per-sequence attention (CPU reference path)."""
import math

import pytest
import torch


def _dense(sq, sk, window, cu):
    off = sk - sq
    q, k = torch.arange(sq)[:, None], torch.arange(sk)[None, :]
    m = k <= q + off
    if window is not None:
        m &= k >= q + off - window[0]
    if cu is not None:
        sid = torch.bucketize(torch.arange(sq), cu[1:], right=True).clamp_(max=len(cu) - 2)
        m &= sid[:, None] == sid[None, :]
    return m


@pytest.mark.parametrize("sq,sk,window,cu", [(300, 300, (64, 0), None), (256, 384, (100, 0), None), (500, 500, None, [0, 130, 131, 400, 480]),
                                             (512, 512, (50, 0), [0, 200, 512]), (64, 64, (-1, 0), [0, 64])])
def test_band_arrays_describe_the_dense_mask(sq, sk, window, cu):
    from megatron_b200 import ops

    cu_t = torch.tensor(cu) if cu is not None else None
    lo, hi = ops.attention_band(sq, sk, window, cu_t)
    q, k = torch.arange(sq)[:, None], torch.arange(sk)[None, :]
    causal = k <= q + sk - sq
    w = window if window is not None and window[0] >= 0 else None
    dense = _dense(sq, sk, w, cu_t)
    assert torch.equal((k >= lo[:, None]) & causal, dense)          # what the forward / dQ kernels mask with
    assert torch.equal((q < hi[None, :]) & causal, dense)           # what the dK/dV kernel masks with
    assert (lo[1:] >= lo[:-1]).all() and (hi[1:] >= hi[:-1]).all()  # monotone: block ranges can be skipped from the first row / last key of a tile
    assert lo.dtype == torch.int32 and hi.dtype == torch.int32
    assert ops.attention_band(sq, sk, None, None) is None and ops.attention_band(sq, sk, (-1, -1), None) is None


def test_packed_sequences_do_not_attend_across_boundaries():
    from megatron_b200.core.packed_seq_params import PackedSeqParams
    from megatron_b200.core.transformer.dot_product_attention import DotProductAttention
    from megatron_b200.core.transformer.enums import AttnMaskType
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=64, num_attention_heads=4, num_query_groups=2, attention_dropout=0.0)
    attn = DotProductAttention(cfg, layer_number=1, attn_mask_type=AttnMaskType.causal, attention_type="self")
    lens = [5, 1, 10]
    t = sum(lens)
    q, k, v = torch.randn(t, 1, 4, 16, requires_grad=True), torch.randn(t, 1, 2, 16, requires_grad=True), torch.randn(t, 1, 2, 16, requires_grad=True)
    cu = torch.tensor([0, 5, 6, 16], dtype=torch.int32)
    psp = PackedSeqParams(qkv_format="thd", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=10, max_seqlen_kv=10)
    out = attn(q, k, v, None, packed_seq_params=psp)
    out.sum().backward()
    g_packed = q.grad.clone()
    q.grad = None
    parts, s0 = [], 0
    for n in lens:
        parts.append(attn(q[s0:s0 + n], k[s0:s0 + n], v[s0:s0 + n], None))
        s0 += n
    sep = torch.cat(parts, 0)
    sep.sum().backward()
    assert torch.allclose(out, sep, atol=1e-5) and torch.allclose(g_packed, q.grad, atol=1e-5)
    plain = attn(q, k, v, None)                                       # without the boundaries the result differs (the old behaviour)
    assert not torch.allclose(plain[6:], out[6:], atol=1e-3)
