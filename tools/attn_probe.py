"""Which library attention backends work for the Llama-3 shape on this box (time + peak memory)."""
import time, torch
s,b,hq,hk,d = 8192,1,32,8,128
q = torch.randn(s,b,hq,d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(s,b,hk,d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(s,b,hk,d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
def bench(name, fn):
    try:
        torch.cuda.reset_peak_memory_stats(); base = torch.cuda.memory_allocated()
        for _ in range(2):
            o = fn(); o.backward(torch.ones_like(o))
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e2=torch.cuda.Event(enable_timing=True)
        e0.record(); o = fn(); e1.record(); o.backward(torch.ones_like(o)); e2.record(); torch.cuda.synchronize()
        fl = 4*s*s*hq*d/2
        print(f"{name}: fwd {e0.elapsed_time(e1):.2f} ms ({fl/e0.elapsed_time(e1)/1e9:.0f} TF)  bwd {e1.elapsed_time(e2):.2f} ms ({2.5*fl/e1.elapsed_time(e2)/1e9:.0f} TF) peak +{(torch.cuda.max_memory_allocated()-base)/2**30:.2f} GiB", flush=True)
    except Exception as e:
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
from torch.nn.attention import sdpa_kernel, SDPBackend
def sdpa(backend):
    def f():
        with sdpa_kernel(backend):
            o = torch.nn.functional.scaled_dot_product_attention(q.permute(1,2,0,3), k.permute(1,2,0,3), v.permute(1,2,0,3), is_causal=True, enable_gqa=True)
        return o.permute(2,0,1,3)
    return f
for n,be in [("sdpa-flash",SDPBackend.FLASH_ATTENTION),("sdpa-cudnn",SDPBackend.CUDNN_ATTENTION),("sdpa-efficient",SDPBackend.EFFICIENT_ATTENTION)]:
    bench(n, sdpa(be))
try:
    from flash_attn import flash_attn_func
    bench("flash_attn2", lambda: flash_attn_func(q.transpose(0,1), k.transpose(0,1), v.transpose(0,1), causal=True).transpose(0,1))
except Exception as e:
    print("flash_attn import failed", e)

# our tcgen05 forward (+ cuDNN backward on its out/LSE) vs the all-library path
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops
for impl, variant in (("native", 0), ("native", 1), ("library", 0)):
    ops.set_attention_impl(impl)
    ops._FA_VARIANT = variant
    bench(f"megatron_b200[{impl}{'' if impl == 'library' else ',P in TMEM' if variant else ',P via smem'}]", lambda: ops.flash_attention(q, k, v, causal=True))

# our tcgen05 backward (first version) next to the cuDNN backward
ops.set_attention_impl("native"); ops._FA_VARIANT = 1; ops._ATTN_BWD_IMPL = "native"
bench("megatron_b200[native fwd + native bwd]", lambda: ops.flash_attention(q, k, v, causal=True))
ops._ATTN_BWD_IMPL = "library"
