"""Small-shape pass over every kernel family for ``compute-sanitizer`` (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck  python tools/sanitize_ops.py
    compute-sanitizer --tool racecheck python tools/sanitize_ops.py --group simple      # shared-memory hazards of the non-tensor-core kernels

Shapes are tiny because the sanitizer serialises and instruments every access; ``--group`` selects kernel families (tcgen05 / TMA kernels are only
meaningful under memcheck: racecheck does not model mbarrier- and tensor-memory-ordered accesses)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops  # noqa: E402


def simple():
    dev = os.environ.get("SANITIZE_DEV", "cuda")
    x = torch.randn(37, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.ones(512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.rms_norm(x, w).sum().backward()
    ops.layer_norm(x, w, torch.zeros_like(w)).sum().backward()
    y, h = ops.add_rms_norm(x, x.detach(), w)
    (y.sum() + h.sum()).backward()
    g = torch.randn(19, 2 * 256, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.swiglu(g).sum().backward()
    ops.quick_geglu(g).sum().backward()
    ops.squared_relu(g).sum().backward()
    t = torch.randn(33, 2, 4, 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    f = torch.randn(33, 32, device=dev)
    ops.apply_rope(t, torch.cat([f, f], -1)[:, None, None, :]).sum().backward()
    ops.apply_rope_thd(t.reshape(66, 4, 64), torch.tensor([0, 20, 66], device=dev, dtype=torch.int32), torch.cat([f, f], -1).repeat(2, 1)).sum().backward()
    lg = torch.randn(45, 1000, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.vocab_parallel_cross_entropy(lg, torch.randint(0, 1000, (45,), device=dev)).sum().backward()
    sm = torch.randn(2, 2, 17, 40, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.scaled_masked_softmax(sm, torch.rand(2, 1, 17, 40, device=dev) > 0.5, 0.3).sum().backward()
    cx = torch.randn(2, 24, 100, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.causal_conv1d(cx, torch.randn(24, 4, device=dev, dtype=torch.bfloat16, requires_grad=True), None, None).sum().backward()
    st = torch.randn(1, 5, 2, 16, 32, device=dev, requires_grad=True)
    p, fin = ops.ssd_state_passing(st, -torch.rand(1, 2, 5, device=dev))
    (p.sum() + fin.sum()).backward()
    q, sf = ops.mxfp8_quantize(torch.randn(40, 256, device=dev).bfloat16())
    ops.mxfp8_dequantize(q, sf)
    ops.nvfp4_quantize(torch.randn(40, 256, device=dev).bfloat16())
    ps_ = [torch.randn(1000, device=dev) for _ in range(3)]
    ops.multi_tensor_l2norm(ps_)
    ops.multi_tensor_scale(ps_, 0.5)
    logits = torch.randn(64, 8, device=dev)
    ops.moe_topk_router(logits, 2, "softmax")
    torch.cuda.synchronize()
    print("simple group ok", flush=True)


def tensor_core():
    dev = os.environ.get("SANITIZE_DEV", "cuda")
    a, b = torch.randn(256, 128, device=dev).bfloat16(), torch.randn(384, 128, device=dev).bfloat16()
    ops.gemm_nt(a, b)
    ops.gemm_nn(torch.randn(256, 384, device=dev).bfloat16(), b)
    ops.gemm_tn(a, torch.randn(256, 64, device=dev).bfloat16())
    (aq, asf), (bq, bsf) = ops.mxfp8_quantize(a), ops.mxfp8_quantize(b)
    ops.gemm_mxfp8_nt(aq, asf, bq, bsf)
    a4, b4 = torch.randn(256, 256, device=dev).bfloat16(), torch.randn(384, 256, device=dev).bfloat16()
    ops.gemm_nvfp4_nt(*ops.nvfp4_quantize(a4), *ops.nvfp4_quantize(b4))
    q = torch.randn(256, 1, 4, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(256, 1, 2, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(256, 1, 2, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.set_attention_impl("native")
    os.environ["MEGATRON_B200_ATTN_BWD"] = "native"
    ops.flash_attention(q, k, v, causal=True).sum().backward()
    torch.cuda.synchronize()
    print("tensor-core group ok", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="all", choices=["all", "simple", "tensor_core"])
    a = ap.parse_args()
    if a.group in ("all", "simple"):
        simple()
    if a.group in ("all", "tensor_core"):
        tensor_core()
