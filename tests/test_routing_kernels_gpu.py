"""routing_kernels.cu (indices <-> multi-hot, pad routing map, fused aux loss, MLA rotary in place / kv split) vs the PyTorch formulations."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cpu(fn, *a):
    return fn(*[x.cpu() if isinstance(x, torch.Tensor) else x for x in a])


@pytest.mark.parametrize("T,k,E", [(1000, 2, 8), (4096, 8, 256), (77, 6, 40), (5, 1, 3)])
def test_indices_multihot_roundtrip(T, k, E):
    from megatron_b200.core.fusions.fused_indices_converter import fused_indices_to_multihot, fused_multihot_to_indices

    torch.manual_seed(0)
    idx = torch.stack([torch.randperm(E)[:k] for _ in range(T)]).cuda()
    idx[torch.rand(T, k, device="cuda") < 0.1] = -1                                     # dropped slots
    probs = torch.rand(T, k, device="cuda", requires_grad=True)
    m, p = fused_indices_to_multihot(idx, probs, E)
    m_ref, p_ref = _cpu(fused_indices_to_multihot, idx, probs.detach(), E)
    assert m.dtype == torch.bool and torch.equal(m.cpu(), m_ref) and torch.equal(p.detach().cpu(), p_ref)
    w = torch.rand(T, E, device="cuda")
    (p * w).sum().backward()
    want = torch.where(idx >= 0, w.gather(1, idx.clamp(min=0)), torch.zeros_like(probs))
    assert torch.equal(probs.grad, want)
    # inverse: expert order, -1 padded
    pe = p.detach().clone().requires_grad_(True)
    i2, p2 = fused_multihot_to_indices(m, pe, k)
    i2_ref, p2_ref = _cpu(fused_multihot_to_indices, m, pe.detach(), k)
    assert torch.equal(i2.cpu(), i2_ref) and torch.equal(p2.detach().cpu(), p2_ref)
    g = torch.rand(T, k, device="cuda")
    (p2 * g).sum().backward()
    want2 = torch.zeros(T, E, device="cuda")
    valid = i2 >= 0
    rows = torch.arange(T, device="cuda").unsqueeze(1).expand(T, k)[valid]
    want2[rows, i2[valid]] = g[valid]
    assert torch.equal(pe.grad, want2)


@pytest.mark.parametrize("T,E,mult", [(4096, 8, 16), (1000, 64, 128), (300, 5, 7), (17, 3, 32)])
def test_pad_routing_map(T, E, mult):
    from megatron_b200.core.fusions.fused_pad_routing_map import fused_pad_routing_map

    torch.manual_seed(1)
    rm = (torch.rand(T, E, device="cuda") < 0.2)
    out = fused_pad_routing_map(rm, mult)
    ref = fused_pad_routing_map(rm.cpu(), mult)
    assert torch.equal(out.cpu(), ref)
    full = out.sum(0) % mult == 0
    could = (~rm).sum(0) >= (mult - rm.sum(0) % mult) % mult                            # columns with enough zeros to pad
    assert bool((full | ~could).all())


def test_fused_aux_loss_matches_eager():
    from megatron_b200.core.transformer.moe.moe_utils import switch_load_balancing_loss_func

    torch.manual_seed(2)
    T, E, k = 8192, 64, 6
    probs = torch.softmax(torch.randn(T, E, device="cuda"), -1).requires_grad_(True)
    tpe = torch.randint(0, 2 * T * k // E, (E,), device="cuda").float()
    a = switch_load_balancing_loss_func(probs, tpe, T, k, E, 1e-2, fused=True)
    (a * 3.0).backward()
    ga = probs.grad.clone()
    probs.grad = None
    b = switch_load_balancing_loss_func(probs, tpe, T, k, E, 1e-2)
    (b * 3.0).backward()
    assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item()) and torch.allclose(ga, probs.grad, rtol=1e-6, atol=0)
    a2 = switch_load_balancing_loss_func(probs, tpe, T, k, E, 1e-2, fused=True)
    assert a2.item() == a.item()                                                         # deterministic reduction order


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("interleaved", [False, True])
def test_mla_rope_q_inplace_and_kv_split(dtype, interleaved):
    from megatron_b200.core.fusions import fused_mla_yarn_rope_apply as F

    torch.manual_seed(3)
    s, b, n, nope, emb, vd = 96, 2, 16, 128, 64, 128
    ang = (torch.rand(s, 1, 1, emb // 2, device="cuda") * 6.0).repeat(1, 1, 1, 2)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    q0 = torch.randn(s, b, n, nope + emb, device="cuda", dtype=dtype)
    q = q0.clone().requires_grad_(True)
    out = F.fused_apply_mla_rope_for_q(q, ang, nope, emb, 1.2, interleaved)
    tmp = q.detach().clone()
    out_ip = F.fused_apply_mla_rope_for_q(tmp, ang, nope, emb, 1.2, interleaved, inplace=True)
    assert out_ip.data_ptr() == tmp.data_ptr() and torch.equal(out_ip, out.detach()) and torch.equal(q.detach(), q0)
    qr = q0.float().cpu().requires_grad_(True)
    ref = F.fused_apply_mla_rope_for_q(qr, ang.cpu(), nope, emb, 1.2, interleaved)
    assert (out.float().cpu() - ref).abs().max().item() < tol * 4
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g.float().cpu())
    assert (q.grad.float().cpu() - qr.grad).abs().max().item() < tol * 4
    # packed (THD) positions
    pos = torch.randint(0, s, (s * b,), device="cuda")
    qt = q0.reshape(s * b, n, nope + emb).clone()
    o2 = F.fused_apply_mla_rope_for_q(qt.clone(), ang, nope, emb, 1.0, interleaved, position_ids=pos)
    r2 = F.fused_apply_mla_rope_for_q(qt.float().cpu(), ang.cpu(), nope, emb, 1.0, interleaved, position_ids=pos.cpu())
    assert (o2.float().cpu() - r2).abs().max().item() < tol * 4

    kv = torch.randn(s, b, n, nope + vd, device="cuda", dtype=dtype, requires_grad=True)
    kpe = torch.randn(s, b, 1, emb, device="cuda", dtype=dtype, requires_grad=True)
    for angles in (ang, None):
        kv.grad = kpe.grad = None
        key, val = F.fused_apply_mla_rope_for_kv(kv, kpe, angles, emb, nope, vd, 1.2, interleaved)
        kvr, kper = kv.detach().float().cpu().requires_grad_(True), kpe.detach().float().cpu().requires_grad_(True)
        keyr, valr = F.fused_apply_mla_rope_for_kv(kvr, kper, None if angles is None else angles.cpu(), emb, nope, vd, 1.2, interleaved)
        assert key.shape == (s, b, n, nope + emb) and val.shape == (s, b, n, vd)
        assert (key.float().cpu() - keyr).abs().max().item() < tol * 4 and torch.equal(val.float().cpu(), valr)
        gk, gv = torch.randn_like(key), torch.randn_like(val)
        (key * gk).sum().backward(retain_graph=True)
        (val * gv).sum().backward()
        (keyr * gk.float().cpu()).sum().backward(retain_graph=True)
        (valr * gv.float().cpu()).sum().backward()
        assert (kv.grad.float().cpu() - kvr.grad).abs().max().item() < tol * 4
        assert (kpe.grad.float().cpu() - kper.grad).abs().max().item() < tol * 16        # a sum over 16 heads in the storage dtype


def test_mla_attention_uses_the_fused_rope_kernels():
    """MLASelfAttention forward + backward on the GPU equals the same module with the fusion switched off."""
    import torch.distributed as dist

    from megatron_b200 import ops
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.spec_utils import build_module
    from megatron_b200.core.transformer.transformer_config import MLATransformerConfig
    import os

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29733")
        dist.init_process_group("nccl", rank=0, world_size=1)
    mine = not ps.model_parallel_is_initialized()             # another test of this process may have left the groups up
    if mine:
        ps.initialize_model_parallel()
    try:
        model_parallel_cuda_manual_seed(7)
        cfg = MLATransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=8, q_lora_rank=128, kv_lora_rank=128, qk_head_dim=64, qk_pos_emb_head_dim=32,
                                   v_head_dim=64, bf16=True, params_dtype=torch.bfloat16, rope_type="yarn", rotary_interleaved=True, attention_dropout=0.0, hidden_dropout=0.0)
        spec = get_gpt_layer_local_spec(multi_latent_attention=True)
        attn = build_module(spec.submodules.self_attention, config=cfg, layer_number=1).cuda()
        x = torch.randn(256, 2, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        outs = []
        for fused in (True, False):
            attn.fused_rope = fused
            x.grad = None
            ops.reset_launch_count()
            y, _ = attn(x, None)
            y.float().square().mean().backward()
            outs.append((y.detach().float(), x.grad.float().clone(), ops.launch_count()))
        assert (outs[0][0] - outs[1][0]).abs().max().item() < 2e-2 * outs[1][0].abs().max().item()
        assert (outs[0][1] - outs[1][1]).abs().max().item() < 3e-2 * outs[1][1].abs().max().item()
    finally:
        if mine:
            ps.destroy_model_parallel()
