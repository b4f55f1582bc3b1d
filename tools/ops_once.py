"""One call of every memory-bound kernel at the Llama-3 8B shapes (target for ``ncu --set full -k regex:...``)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
S, H, F, V = 8192, 4096, 14336, 128256
dev = "cuda"
x = torch.randn(S, 1, H, device=dev).bfloat16().requires_grad_(True)
w = torch.ones(H, device=dev).bfloat16().requires_grad_(True)
for _ in range(2):
    y = ops.rms_norm(x, w, 1e-5); y.backward(torch.ones_like(y))
    g = torch.randn(S, 1, 2 * F, device=dev).bfloat16().requires_grad_(True)
    z = ops.swiglu(g); z.backward(torch.ones_like(z))
    q = torch.randn(S, 1, 32, 128, device=dev).bfloat16().requires_grad_(True)
    fr = torch.randn(S, 1, 1, 128, device=dev)
    r = ops.apply_rope(q, fr); r.backward(torch.ones_like(r))
    lg = torch.randn(S // 4, 1, V, device=dev).bfloat16().requires_grad_(True)
    tg = torch.randint(0, V, (S // 4, 1), device=dev)
    ce = ops.vocab_parallel_cross_entropy(lg, tg); ce.sum().backward()
    p32 = [torch.randn(64 << 20, device=dev)]; gr = [torch.randn(64 << 20, device=dev).bfloat16()]; m = [torch.zeros_like(p32[0])]; v = [torch.zeros_like(p32[0])]; lo = [p32[0].bfloat16()]
    ops.fused_adam(p32, gr, m, v, lo, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=1)
    n = ops.multi_tensor_l2norm(gr)
torch.cuda.synchronize(); print("done", float(n))
