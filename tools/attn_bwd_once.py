"""Run the native flash-attention backward a few times at the Llama-3 8B shape (target for ``ncu -k regex:fa_bwd``)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
s, b, hq, hk, d = 8192, 1, 32, 8, 128
q = torch.randn(s, b, hq, d, device="cuda", dtype=torch.bfloat16)
k = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
v = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
go = torch.randn(s, b, hq, d, device="cuda", dtype=torch.bfloat16)
scale = 1.0 / math.sqrt(d)
o, lse = ops.ext().flash_attn_fwd(q, k, v, True, scale, 1)
delta = (go.float() * o.float()).sum(-1).permute(1, 2, 0).contiguous()
for _ in range(3):
    dq, dk, dv = ops.ext().flash_attn_bwd(go, q, k, v, lse, delta, True, scale)
torch.cuda.synchronize()
print("done", float(dk.float().abs().mean()))
