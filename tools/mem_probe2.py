import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200.training.engine import TrainEngine
G = 2**30
def rep(tag):
    torch.cuda.synchronize()
    print(f"{tag:40s} alloc {torch.cuda.memory_allocated()/G:7.2f} GiB  peak {torch.cuda.max_memory_allocated()/G:7.2f} GiB", flush=True)
rc = dict(recompute_granularity="selective", recompute_modules=["layernorm", "mlp_act"])
eng = TrainEngine("llama3_8b", micro_batch_size=1, global_batch_size=4, bf16=True, model_overrides=dict(num_layers=4), **rc)
rep("init")
orig = eng._forward_step
cnt = [0]
def fs(it, model):
    rep(f"  before fwd mb{cnt[0]}")
    o = orig(it, model)
    rep(f"  after  fwd mb{cnt[0]}")
    cnt[0] += 1
    return o
eng._forward_step = fs
tok = eng.synthetic_batch().cuda()
for step in range(3):
    torch.cuda.reset_peak_memory_stats()
    cnt[0] = 0
    l = eng.train_step(tok)
    rep(f"step {step} done")
import gc
objs = [o for o in gc.get_objects() if torch.is_tensor(o) and o.is_cuda]
seen = {}
for o in objs:
    try:
        st = o.untyped_storage()
        seen[st.data_ptr()] = max(seen.get(st.data_ptr(), 0), st.nbytes())
    except Exception:
        pass
big = sorted(seen.values(), reverse=True)[:25]
print("live cuda storages (GiB):", [round(b / G, 2) for b in big], "total", round(sum(seen.values()) / G, 2))
