"""Time the native flash-attention backward (and forward) at the Llama-3 8B shape for the head counts of TP=1/2/4/8, next to cuDNN.
Also the target for ``ncu -k regex:fa_bwd``.  CUDA events after warm-up; FLOPs = 4*s^2*h*d/2 (causal) fwd, x2.5 bwd."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
s, b, d = int(os.environ.get("SEQ", "8192")), 1, 128
iters = int(os.environ.get("ITERS", "5"))
scale = 1.0 / math.sqrt(d)

def timeit(fn, n=iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for hq, hk in [(32, 8), (16, 4), (8, 2), (4, 1)][: int(os.environ.get("NCONF", "4"))]:
    q = torch.randn(s, b, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(s, b, hk, d, device="cuda", dtype=torch.bfloat16)
    go = torch.randn(s, b, hq, d, device="cuda", dtype=torch.bfloat16)
    fl = 4 * s * s * hq * d / 2
    o, lse = ops.ext().flash_attn_fwd(q, k, v, True, scale, 1)
    t_f = timeit(lambda: ops.ext().flash_attn_fwd(q, k, v, True, scale, 1))
    res = {}
    for sh in (0, 1):
        res[sh] = timeit(lambda: ops.ext().flash_attn_bwd(go, q, k, v, o, lse, True, scale, sh))
    lse_lib = lse.unsqueeze(-1) if ops._cudnn_lse_ndim() == 4 else lse
    qb, kb, vb = (t.permute(1, 2, 0, 3) for t in (q, k, v))
    from torch.nn.attention import SDPBackend, sdpa_kernel
    def lib_fwd():
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
            return torch.nn.functional.scaled_dot_product_attention(qb, kb, vb, is_causal=True, scale=scale, enable_gqa=True)
    try:
        t_lf = timeit(lib_fwd)
    except Exception as e:
        t_lf = float("nan")
    os.environ["MB200_FA_PAIR_MODE"] = "0"
    t_f0 = timeit(lambda: ops.ext().flash_attn_fwd(q, k, v, True, scale, 1))
    os.environ["MB200_FA_PAIR_MODE"] = "1"
    t_f1 = timeit(lambda: ops.ext().flash_attn_fwd(q, k, v, True, scale, 1))
    del os.environ["MB200_FA_PAIR_MODE"]
    print(f"heads {hq}/{hk}: fwd neighbours {t_f0:.3f} ms ({fl/t_f0/1e9:.0f} TF), mirrored {t_f1:.3f} ms ({fl/t_f1/1e9:.0f} TF), auto {t_f:.3f} ms | cuDNN fwd {t_lf:.3f} ms ({fl/t_lf/1e9:.0f} TF)", flush=True)
    t_lib = timeit(lambda: ops.ext().attn_bwd_cudnn(go, q, k, v, o, lse_lib, True, scale))
    print(f"heads {hq}/{hk}: fwd ours {t_f:.3f} ms ({fl/t_f/1e9:.0f} TF) | bwd ours fused-heads {res[0]:.3f} ms ({2.5*fl/res[0]/1e9:.0f} TF), split-heads {res[1]:.3f} ms "
          f"({2.5*fl/res[1]/1e9:.0f} TF) | cuDNN bwd {t_lib:.3f} ms ({2.5*fl/t_lib/1e9:.0f} TF)", flush=True)
    if os.environ.get("CHECK", "1") == "1" and hq == 32:
        dq, dk, dv = ops.ext().flash_attn_bwd(go, q, k, v, o, lse, True, scale, 0)
        rq, rk, rv = ops.ext().attn_bwd_cudnn(go, q, k, v, o, lse_lib, True, scale)
        for nm, a, r in zip("qkv", (dq, dk, dv), (rq, rk, rv)):
            print(f"   d{nm} vs cuDNN: max rel err {float((a.float()-r.float()).abs().max()/r.float().abs().max()):.4f}", flush=True)
