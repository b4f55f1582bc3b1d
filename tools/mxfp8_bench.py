"""Throughput of the block-scaled MXFP8 GEMM vs per-tensor FP8 and bf16 (CUDA events, L2-sized operands rotated between iterations)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops  # noqa: E402


def bench(fn, iters=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    for M, N, K in [(8192, 8192, 8192), (8192, 28672, 4096), (8192, 4096, 14336)]:
        nbuf = 3   # rotate operand sets so consecutive iterations do not hit the same lines in L2
        A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
        B = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(nbuf)]
        q = [(ops.mxfp8_quantize(a), ops.mxfp8_quantize(b)) for a, b in zip(A, B)]
        sw = [((aq, ops.mxfp8_swizzle_scales(asf)), (bq, ops.mxfp8_swizzle_scales(bsf))) for (aq, asf), (bq, bsf) in q]
        ext = ops.ext()
        t_mx = bench(lambda i: ext.gemm_mxfp8_nt(sw[i % nbuf][0][0], sw[i % nbuf][0][1], sw[i % nbuf][1][0], sw[i % nbuf][1][1], 256))
        t_mx128 = bench(lambda i: ext.gemm_mxfp8_nt(sw[i % nbuf][0][0], sw[i % nbuf][0][1], sw[i % nbuf][1][0], sw[i % nbuf][1][1], 128))
        f8 = [(a.to(torch.float8_e4m3fn), b.to(torch.float8_e4m3fn)) for a, b in zip(A, B)]
        t_f8 = bench(lambda i: ext.gemm_fp8_nt(f8[i % nbuf][0], f8[i % nbuf][1], 1.0, None))
        t_bf = bench(lambda i: ops.gemm_nt(A[i % nbuf], B[i % nbuf]))
        t_q = bench(lambda i: ops.mxfp8_quantize(A[i % nbuf]))
        t_f4 = float("nan")
        if hasattr(ext, "gemm_nvfp4_nt") and K % 256 == 0:
            from megatron_b200.core.fp4_utils import quantize_nvfp4

            f4 = []
            for a, b in zip(A, B):
                (ac, asc, _), (bc, bsc, _) = quantize_nvfp4(a[:, :K]), quantize_nvfp4(b[:, :K])
                f4.append((ops.nvfp4_pack(ac), ops.mxfp8_swizzle_scales(asc.view(torch.uint8)), ops.nvfp4_pack(bc), ops.mxfp8_swizzle_scales(bsc.view(torch.uint8))))
                del ac, asc, bc, bsc
            t_f4 = bench(lambda i: ext.gemm_nvfp4_nt(f4[i % nbuf][0], f4[i % nbuf][1], f4[i % nbuf][2], f4[i % nbuf][3], 1.0, None))
        fl = 2.0 * M * N * K
        print(f"M{M} N{N} K{K}: mxfp8[128x256] {t_mx:.3f} ms ({fl / t_mx / 1e9:.0f} TF) | mxfp8[128x128] {t_mx128:.3f} ms ({fl / t_mx128 / 1e9:.0f} TF) | fp8 per-tensor {t_f8:.3f} ms ({fl / t_f8 / 1e9:.0f} TF) | bf16 {t_bf:.3f} ms ({fl / t_bf / 1e9:.0f} TF)"
              f" | nvfp4 {t_f4:.3f} ms ({fl / t_f4 / 1e9:.0f} TF) | quantise A {t_q * 1e3:.0f} us ({(M * K * 3 + M * K / 32) / t_q / 1e9:.2f} TB/s)", flush=True)


if __name__ == "__main__":
    main()
