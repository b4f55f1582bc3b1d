"""Per-layer CUDA graph capture/replay equals eager execution (fwd values, input grads, param grads)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_module_matches_eager():
    from megatron_b200.core.transformer.cuda_graphs import graph_module

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256)).cuda()
    ref = copy.deepcopy(net)
    net = graph_module(net, warmup_steps=2)
    for it in range(6):
        x = torch.randn(64, 256, device="cuda", requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        y = net(x)
        assert net.cudagraph_manager.fallback_reason is None, net.cudagraph_manager.fallback_reason
        yr = ref(xr)
        gy = torch.randn_like(yr)
        y.backward(gy)
        yr.backward(gy)
        assert torch.allclose(y, yr, atol=1e-5), it
        assert torch.allclose(x.grad, xr.grad, atol=1e-5), it
    assert net.cudagraph_manager.fallback_reason is None
    assert len(net.cudagraph_manager.captured) == 1
    for p, pr in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, pr.grad, atol=1e-4)
