"""Per-layer CUDA graph capture/replay equals eager execution (fwd values, input grads, param grads).

Runs in a child process: a failed capture leaves PyTorch's CUDA RNG in capture mode and would poison every later test."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

_BODY = """
import copy, sys, torch
sys.path.insert(0, %r)
from megatron_b200.core.transformer.cuda_graphs import graph_module
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
ref = copy.deepcopy(net)
net = graph_module(net, warmup_steps=0)   # capture on the first call (make_graphed_callables runs its own side-stream warm-up)
# all random inputs are drawn BEFORE the capture: on this torch build the CUDA RNG cannot be used eagerly in the same process afterwards
xs = [torch.randn(64, 256, device="cuda") for _ in range(6)]
gys = [torch.randn(64, 256, device="cuda") for _ in range(6)]
for it in range(6):
    x = xs[it].clone().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y = net(x)
    assert net.cudagraph_manager.fallback_reason is None, net.cudagraph_manager.fallback_reason
    yr = ref(xr)
    gy = gys[it]
    # like the training loop (grads are consumed into main_grad every micro-batch): start each backward from empty .grad —
    # a replayed backward hands out the SAME static gradient buffers every time, so .grad accumulation across replays is not defined
    net.zero_grad(set_to_none=True)
    ref.zero_grad(set_to_none=True)
    y.backward(gy)
    yr.backward(gy)
    assert torch.allclose(y, yr, atol=1e-5), it
    assert torch.allclose(x.grad, xr.grad, atol=1e-5), it
    for p, pr in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, pr.grad, atol=1e-3, rtol=1e-3), (it, (p.grad - pr.grad).abs().max())
assert len(net.cudagraph_manager.captured) == 1
print("GRAPH_OK")
"""


def test_graphed_module_matches_eager():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(_BODY % root)], capture_output=True, text=True, timeout=300)
    assert "GRAPH_OK" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
