import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch.distributed as dist
from megatron_b200 import ops
from megatron_b200.core.fusions import fused_mla_yarn_rope_apply as F
torch.manual_seed(3)
# kernel level, emb = 32
s, b, n, nope, emb, vd = 64, 2, 8, 64, 32, 64
ang = (torch.rand(s, 1, 1, emb // 2, device="cuda") * 6.0).repeat(1, 1, 1, 2)
for il in (False, True):
    for dt in (torch.float32, torch.bfloat16):
        q0 = torch.randn(s, b, n, nope + emb, device="cuda", dtype=dt)
        q = q0.clone().requires_grad_(True)
        out = F.fused_apply_mla_rope_for_q(q, ang, nope, emb, 1.2, il)
        qr = q0.float().cpu().requires_grad_(True)
        ref = F.fused_apply_mla_rope_for_q(qr, ang.cpu(), nope, emb, 1.2, il)
        g = torch.randn_like(out); out.backward(g); ref.backward(g.float().cpu())
        print("q", il, dt, (out.float().cpu() - ref).abs().max().item(), (q.grad.float().cpu() - qr.grad).abs().max().item())
        kv = torch.randn(s, b, n, nope + vd, device="cuda", dtype=dt, requires_grad=True); kpe = torch.randn(s, b, 1, emb, device="cuda", dtype=dt, requires_grad=True)
        key, val = F.fused_apply_mla_rope_for_kv(kv, kpe, None, emb, nope, vd)
        kvr, kper = kv.detach().float().cpu().requires_grad_(True), kpe.detach().float().cpu().requires_grad_(True)
        keyr, valr = F.fused_apply_mla_rope_for_kv(kvr, kper, None, emb, nope, vd)
        gk, gv = torch.randn_like(key), torch.randn_like(val)
        torch.autograd.backward([key, val], [gk, gv]); torch.autograd.backward([keyr, valr], [gk.float().cpu(), gv.float().cpu()])
        print("kv", il, dt, (key.float().cpu() - keyr).abs().max().item(), (kv.grad.float().cpu() - kvr.grad).abs().max().item(), (kpe.grad.float().cpu() - kper.grad).abs().max().item())
# module level
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29733")
dist.init_process_group("nccl", rank=0, world_size=1)
from megatron_b200.core import parallel_state as ps
from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
from megatron_b200.core.transformer.spec_utils import build_module
from megatron_b200.core.transformer.transformer_config import MLATransformerConfig
ps.initialize_model_parallel(); model_parallel_cuda_manual_seed(7)
for dt in (torch.float32, torch.bfloat16):
    cfg = MLATransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=8, q_lora_rank=128, kv_lora_rank=128, qk_head_dim=64, qk_pos_emb_head_dim=32,
                               v_head_dim=64, bf16=dt == torch.bfloat16, params_dtype=dt, rope_type="yarn", rotary_interleaved=True, attention_dropout=0.0, hidden_dropout=0.0)
    attn = build_module(get_gpt_layer_local_spec(multi_latent_attention=True).submodules.self_attention, config=cfg, layer_number=1).cuda()
    x = torch.randn(256, 2, 512, device="cuda", dtype=dt, requires_grad=True)
    res = []
    for fused in (True, False, True, False):
        attn.fused_rope = fused; x.grad = None
        for p_ in attn.parameters(): p_.grad = None
        y, _ = attn(x, None); y.float().square().mean().backward()
        res.append((y.detach().float(), x.grad.float().clone(), [p_.grad.float().clone() if p_.grad is not None else (p_.main_grad.float().clone() if hasattr(p_, "main_grad") else None) for p_ in attn.parameters()]))
    rel = lambda a, c: ((a - c).abs().max() / (c.abs().max() + 1e-12)).item()
    print(dt, "y fused-vs-unfused", rel(res[0][0], res[1][0]), "dx", rel(res[0][1], res[1][1]), "| repeat fused", rel(res[0][1], res[2][1]), "repeat unfused", rel(res[1][1], res[3][1]))
    for (nme, _), a, c in zip(attn.named_parameters(), res[0][2], res[1][2]):
        if a is not None: print("   dparam", nme, rel(a, c))
