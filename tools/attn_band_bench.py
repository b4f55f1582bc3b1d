"""Band-mask attention (sliding window / packed sequences) inside the native tcgen05 kernels: time fwd and fwd+bwd against the dense causal call on the
same tensors, and report the visited-pairs ratio (the ideal speed-up).  1 GPU."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    s, hq, hk, d = 8192, 32, 8, 128
    q = torch.randn(s, 1, hq, d, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(s, 1, hk, d, device="cuda").bfloat16().requires_grad_(True)
    go = torch.randn(s, 1, hq, d, device="cuda").bfloat16()
    cases = [("dense_causal", None, None), ("window_4096", (4095, 0), None), ("window_1024", (1023, 0), None),
             ("packed_8x1024", None, torch.arange(0, s + 1, 1024, device="cuda", dtype=torch.int32)),
             ("packed_ragged", None, torch.tensor([0, 700, 2100, 2101, 5000, 8192], device="cuda", dtype=torch.int32))]
    dense_pairs = s * (s + 1) / 2
    for name, w, cu in cases:
        band = ops.attention_band(s, s, w, cu, q.device)
        pairs = dense_pairs if band is None else float((torch.arange(s, device="cuda") + 1 - band[0]).sum())
        fwd = lambda: ops.flash_attention(q, k, v, causal=True, window=w, cu_seqlens=cu)
        def fb():
            o = ops.flash_attention(q, k, v, causal=True, window=w, cu_seqlens=cu)
            o.backward(go)
        t_f, t_fb = timeit(fwd), timeit(fb)
        fl = 4 * pairs * hq * d
        print(json.dumps({"bench": "attn_band", "case": name, "s": s, "heads": f"{hq}/{hk}", "pairs_vs_dense": round(pairs / dense_pairs, 3), "fwd_ms": round(t_f, 3),
                          "fwd_TF": round(fl / t_f / 1e9, 0), "fwd_bwd_ms": round(t_fb, 3), "bwd_ms": round(t_fb - t_f, 3), "bwd_TF": round(2.5 * fl / (t_fb - t_f) / 1e9, 0)}), flush=True)

if __name__ == "__main__":
    main()
