"""One forward + backward of the native attention under a band mask (default: sliding window of 1024 keys at s = 8192, 32/8 heads) — the target of
``ncu -k regex:fa_`` for the band-mask code paths (MASK=window|packed|dense)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
s, hq, hk, d = int(os.environ.get("SEQ", "8192")), 32, 8, 128
kind = os.environ.get("MASK", "window")
q = torch.randn(s, 1, hq, d, device="cuda").bfloat16().requires_grad_(True)
k = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
v = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
go = torch.randn(s, 1, hq, d, device="cuda").bfloat16()
w = (1023, 0) if kind == "window" else None
cu = torch.arange(0, s + 1, 1024, device="cuda", dtype=torch.int32) if kind == "packed" else None
for _ in range(int(os.environ.get("ITERS", "3"))):
    o = ops.flash_attention(q, k, v, causal=True, window=w, cu_seqlens=cu)
    o.backward(go)
torch.cuda.synchronize()
print("done", kind)
