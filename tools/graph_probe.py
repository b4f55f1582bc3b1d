"""Which CUDA-graph capture paths work on this box (each attempt in its own process: a failed capture poisons the RNG state)."""
import subprocess
import sys
import textwrap

CASES = {
    "torch.make_graphed_callables(Sequential)": """
        import torch
        net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
        x = torch.randn(64, 256, device="cuda", requires_grad=True)
        g = torch.cuda.make_graphed_callables(net, (x,))
        y = g(x); y.sum().backward(); torch.cuda.synchronize(); print("OK", float(y.sum()))
    """,
    "direct after eager fwd+bwd on the default stream": """
        import torch
        net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
        x = torch.randn(64, 256, device="cuda", requires_grad=True)
        net(x).sum().backward(); torch.cuda.synchronize()
        g = torch.cuda.make_graphed_callables(net, (x.detach().clone().requires_grad_(True),))
        y = g(x); y.sum().backward(); torch.cuda.synchronize(); z = torch.randn(4, device="cuda"); print("OK", float(y.sum()))
    """,
    "direct with pool=graph_pool_handle() + randn afterwards": """
        import torch
        net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
        x = torch.randn(64, 256, device="cuda", requires_grad=True)
        g = torch.cuda.make_graphed_callables(net, (x,), pool=torch.cuda.graph_pool_handle())
        y = g(x); y.sum().backward(); torch.cuda.synchronize(); z = torch.randn(4, device="cuda"); print("OK", float(y.sum()))
    """,
    "direct on a tuple-returning shim module": """
        import torch
        class Shim(torch.nn.Module):
            def __init__(s, inner): super().__init__(); s.inner = inner
            def forward(s, *xs): return (s.inner(*xs),)
        net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
        x = torch.randn(64, 256, device="cuda", requires_grad=True)
        g = torch.cuda.make_graphed_callables(Shim(net), (x,))
        y = g(x)[0]; y.sum().backward(); torch.cuda.synchronize(); print("OK", float(y.sum()))
    """,
    "manager(graph_module), warmup 0": """
        import torch, sys, os
        sys.path.insert(0, os.getcwd())
        from megatron_b200.core.transformer.cuda_graphs import graph_module
        net = graph_module(torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda(), warmup_steps=0)
        xs = [torch.randn(64, 256, device="cuda", requires_grad=True) for _ in range(4)]
        for x in xs:
            y = net(x); y.sum().backward()
        torch.cuda.synchronize(); print("OK" if net.cudagraph_manager.fallback_reason is None else "FALLBACK " + net.cudagraph_manager.fallback_reason[:300], len(net.cudagraph_manager.captured))
    """,
    "manager(graph_module)": """
        import torch, sys, os
        sys.path.insert(0, os.getcwd())
        from megatron_b200.core.transformer.cuda_graphs import graph_module
        net = graph_module(torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda(), warmup_steps=2)
        for it in range(5):
            x = torch.randn(64, 256, device="cuda", requires_grad=True)
            y = net(x); y.sum().backward()
        torch.cuda.synchronize(); print("OK" if net.cudagraph_manager.fallback_reason is None else "FALLBACK " + net.cudagraph_manager.fallback_reason[:300], len(net.cudagraph_manager.captured))
    """,
    "manager(TransformerLayer via TrainEngine tiny_llama, enable_cuda_graph)": """
        import torch, sys, os
        sys.path.insert(0, os.getcwd())
        from megatron_b200.training.engine import TrainEngine
        eng = TrainEngine("tiny_llama", micro_batch_size=2, global_batch_size=2, model_overrides={"enable_cuda_graph": True})
        batch = eng.synthetic_batch()
        for _ in range(6):
            out = float(eng.train_step(batch))
        layer = eng.model[0].module.decoder.layers[0] if hasattr(eng.model[0], "module") else eng.model[0].decoder.layers[0]
        m = layer.cudagraph_manager
        print("OK" if (m is not None and m.fallback_reason is None and len(m.captured) > 0) else f"NOT GRAPHED {getattr(m, 'fallback_reason', None)}", out)
    """,
}
for name, code in CASES.items():
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=300)
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = [ln for ln in r.stderr.splitlines() if "Error" in ln or "error" in ln][-3:]
    print(f"[{name}] rc={r.returncode} :: {tail} :: {' | '.join(err)[:600]}", flush=True)
