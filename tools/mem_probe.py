"""Where does the memory go?  4-layer Llama-3-8B-shaped model, one micro-batch, stage by stage."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200.training.engine import TrainEngine
G = 2**30
def rep(tag):
    torch.cuda.synchronize()
    print(f"{tag:40s} alloc {torch.cuda.memory_allocated()/G:7.2f} GiB  peak {torch.cuda.max_memory_allocated()/G:7.2f} GiB  reserved {torch.cuda.memory_reserved()/G:7.2f}", flush=True)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rc = dict(recompute_granularity="selective", recompute_modules=["layernorm", "mlp_act"]) if "norc" not in sys.argv else {}
eng = TrainEngine("llama3_8b", micro_batch_size=1, global_batch_size=1, bf16=True, model_overrides=dict(num_layers=L), **rc)
n = sum(p.numel() for p in eng.model_chunks[0].parameters())
print("params", n/1e9, "B")
rep("after engine init")
tok = eng.synthetic_batch().cuda()
m = eng.model[0]
pos = torch.arange(8192, device="cuda")[None]
torch.cuda.reset_peak_memory_stats()
base = torch.cuda.memory_allocated()
# layer-by-layer activation accounting with forward hooks
marks = []
def mk(name):
    def hook(mod, inp, out):
        marks.append((name, torch.cuda.memory_allocated()))
    return hook
gpt = eng.model_chunks[0]
for i, layer in enumerate(gpt.decoder.layers):
    layer.register_forward_hook(mk(f"layer{i}"))
    layer.self_attention.register_forward_hook(mk(f"layer{i}.attn"))
    layer.self_attention.linear_qkv.register_forward_hook(mk(f"layer{i}.qkv"))
    layer.self_attention.core_attention.register_forward_hook(mk(f"layer{i}.core"))
    layer.mlp.register_forward_hook(mk(f"layer{i}.mlp"))
gpt.embedding.register_forward_hook(mk("embedding"))
gpt.output_layer.register_forward_hook(mk("output_layer"))
loss = m(tok[:, :-1].contiguous(), pos, None, labels=tok[:, 1:].contiguous())
rep("after forward (1 microbatch)")
prev = base
for name, a in marks:
    print(f"   {name:24s} +{(a-prev)/2**20:9.1f} MiB (cum {(a-base)/G:6.2f} GiB)")
    prev = a
loss.float().mean().backward()
rep("after backward")
eng.optimizer.step()
rep("after optimizer step (states allocated)")
for m_ in eng.model: m_.zero_grad_buffer()
torch.cuda.reset_peak_memory_stats()
l = eng.train_step(tok)
rep("after 2nd full train_step")
