"""Run the native flash-attention forward a few times at the Llama-3 8B shape (target for ``ncu -k regex:fa_fwd``)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
s, b, hq, hk, d = 8192, 1, 32, 8, 128
q = torch.randn(s, b, hq, d, device="cuda", dtype=torch.bfloat16)
k = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
v = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    o, lse = ops.ext().flash_attn_fwd(q, k, v, True, 1.0 / math.sqrt(d))
torch.cuda.synchronize()
print("done", float(o.float().abs().mean()))
