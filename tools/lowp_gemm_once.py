"""Launch the MXFP8 and NVFP4 GEMMs a few times (for ncu capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops  # noqa: E402

M = N = K = 8192
a, b = torch.randn(M, K, device="cuda").bfloat16(), torch.randn(N, K, device="cuda").bfloat16()
(aq, asf), (bq, bsf) = ops.mxfp8_quantize(a), ops.mxfp8_quantize(b)
sa, sb = ops.mxfp8_swizzle_scales(asf), ops.mxfp8_swizzle_scales(bsf)
qa, qb = ops.nvfp4_quantize(a), ops.nvfp4_quantize(b)
fa, fb = ops.mxfp8_swizzle_scales(qa[1].view(torch.uint8)), ops.mxfp8_swizzle_scales(qb[1].view(torch.uint8))
for _ in range(4):
    ops.ext().gemm_mxfp8_nt(aq, sa, bq, sb, 256)
    ops.ext().gemm_nvfp4_nt(qa[0], fa, qb[0], fb, 1.0, None)
torch.cuda.synchronize()
print("done")
