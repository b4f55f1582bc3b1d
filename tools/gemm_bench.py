"""Per-shape TFLOP/s of the tcgen05 GEMM vs cuBLAS for the Llama-3-8B shapes (all three layouts)."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
S = 8192
shapes = {  # name: (M, N, K) of the forward GEMM  y[M,N] = x[M,K] w[N,K]^T
    "qkv_tp1": (S, 6144, 4096), "proj_tp1": (S, 4096, 4096), "fc1_tp1": (S, 28672, 4096), "fc2_tp1": (S, 4096, 14336), "lmhead_tp1": (S, 128256, 4096),
    "qkv_tp8": (S, 768, 4096), "proj_tp8": (S, 4096, 512), "fc1_tp8": (S, 3584, 4096), "fc2_tp8": (S, 4096, 1792),
}
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
rows = []
only = sys.argv[1:] if len(sys.argv) > 1 else list(shapes)
for name in only:
    M, N, K = shapes[name]
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); gy = torch.randn(M, N, device="cuda").bfloat16()
    fl = 2.0 * M * N * K
    res = {"shape": name, "MNK": [M, N, K]}
    for be in ("tcgen05:3", "tcgen05:4", "tcgen05:5", "tcgen05:6", "cublas"):
        ops.set_gemm_backend(be)
        res[f"fwd_{be}"] = fl / timeit(lambda: ops.gemm_nt(x, w)) / 1e9
        res[f"dgrad_{be}"] = fl / timeit(lambda: ops.gemm_nn(gy, w)) / 1e9
        res[f"wgrad_{be}"] = fl / timeit(lambda: ops.gemm_tn(gy, x)) / 1e9
    rows.append(res)
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    del x, w, gy
