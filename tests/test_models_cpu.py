"""Model families and attention variants beyond GPT: MLA, MTP, BERT, T5, hybrid Mamba, LLaVA, FSDP (CPU, gloo)."""
import torch

from dist_utils import run_distributed


def _init(seed=1, tp=1):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(tp, 1)
    model_parallel_cuda_manual_seed(seed)


_KW = dict(use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)


def _incremental_matches_full(m, tok, pos, prefill):
    from megatron_b200.core.inference_params import InferenceParams

    m.eval()
    with torch.no_grad():
        full = m(tok, pos, None)
        ip = InferenceParams(tok.shape[0], 64)
        outs = [m(tok[:, :prefill], pos[:, :prefill], None, inference_context=ip)]
        ip.sequence_len_offset = prefill
        for t in range(prefill, tok.shape[1]):
            outs.append(m(tok[:, t : t + 1], pos[:, t : t + 1], None, inference_context=ip))
            ip.sequence_len_offset += 1
    return (torch.cat(outs, 1) - full).abs().max().item()


def _mla(rank, world):
    _init()
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import MLATransformerConfig

    cfg = MLATransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, q_lora_rank=16, kv_lora_rank=24, qk_head_dim=16,
                               qk_pos_emb_head_dim=8, v_head_dim=16, ffn_hidden_size=128, gated_linear_unit=True,
                               activation_func=torch.nn.functional.silu, add_bias_linear=False, rotary_scaling_factor=4.0,
                               original_max_position_embeddings=16, mscale_all_dim=1.0, qk_layernorm=True, **_KW)
    spec = get_gpt_layer_local_spec(multi_latent_attention=True, qk_layernorm=True, normalization="RMSNorm")
    m = GPTModel(cfg, spec, vocab_size=128, max_sequence_length=64, position_embedding_type="none")
    tok = torch.randint(0, 128, (2, 32))
    pos = torch.arange(32)[None].expand(2, -1)
    m(tok, pos, None, labels=tok).mean().backward()
    assert all(p.grad is not None for p in m.parameters())
    # the latent KV cache (kv_lora_rank + rope dims per token) reproduces the full forward
    assert _incremental_matches_full(m, tok, pos, 20) < 1e-5
    return True


def test_mla_trains_and_latent_cache_decodes():
    run_distributed(_mla, 1)


def _mtp(rank, world):
    _init()
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec, get_gpt_mtp_block_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.multi_token_prediction import MTPLossLoggingHelper, roll_tensor
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    r, s = roll_tensor(torch.arange(1, 7).view(1, 6).float(), -1, -1)
    assert r.tolist() == [[2, 3, 4, 5, 6, 0]] and s.item() == 20
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, ffn_hidden_size=128, gated_linear_unit=True,
                            activation_func=torch.nn.functional.silu, add_bias_linear=False, normalization="RMSNorm", mtp_num_layers=2,
                            mtp_loss_scaling_factor=0.3, **_KW)
    spec = get_gpt_layer_local_spec(normalization="RMSNorm")
    m = GPTModel(cfg, spec, vocab_size=128, max_sequence_length=64, position_embedding_type="rope", mtp_block_spec=get_gpt_mtp_block_spec(cfg, spec))
    tok = torch.randint(0, 128, (2, 32))
    pos = torch.arange(32)[None].expand(2, -1)
    m(tok, pos, None, labels=tok, loss_mask=torch.ones(2, 32)).mean().backward()
    assert all(p.grad is not None for p in m.parameters()), "MTP losses must reach the MTP layers through the main loss"
    logged = MTPLossLoggingHelper.pop()
    assert set(logged) == {"mtp_1 loss", "mtp_2 loss"} and all(3.0 < v.item() < 7.0 for v in logged.values())
    return True


def test_mtp_losses_backprop_through_main_loss():
    run_distributed(_mtp, 1)


def _bert_t5(rank, world):
    _init()
    from megatron_b200.core.models.bert.bert_layer_specs import bert_layer_local_spec
    from megatron_b200.core.models.bert.bert_model import BertModel
    from megatron_b200.core.models.T5.t5_model import T5Model
    from megatron_b200.core.models.T5.t5_spec import get_t5_decoder_with_local_block_spec, get_t5_encoder_with_local_block_spec
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, **_KW)
    m = BertModel(cfg, 2, bert_layer_local_spec, 128, 64, share_embeddings_and_output_weights=True)
    tok = torch.randint(0, 128, (2, 16))
    am = torch.ones(2, 16)
    am[1, 10:] = 0
    loss, bl = m(tok, am, tokentype_ids=torch.zeros_like(tok), lm_labels=tok)
    (loss.mean() + bl.sum()).backward()
    m.eval()
    with torch.no_grad():
        l1, _ = m(tok, am)
        tok2 = tok.clone()
        tok2[1, 10:] = 5
        l2, _ = m(tok2, am)
    assert (l1[1, :10] - l2[1, :10]).abs().max().item() < 1e-6, "padded tokens must not influence real tokens"
    t5 = T5Model(cfg, cfg, get_t5_encoder_with_local_block_spec(2), get_t5_decoder_with_local_block_spec(2), 128, 64, share_embeddings_and_output_weights=True)
    enc, dec = torch.randint(0, 128, (2, 16)), torch.randint(0, 128, (2, 12))
    em, dm, xm = torch.ones(2, 16, 16), torch.ones(2, 12, 12), torch.ones(2, 12, 16)
    t5(enc, dec, em, dm, xm, lm_labels=dec).mean().backward()
    assert all(p.grad is not None for p in t5.parameters())
    t5.eval()
    with torch.no_grad():
        a = t5(enc, dec, em, dm, xm)
        dec2 = dec.clone()
        dec2[:, 8:] = 3
        b = t5(enc, dec2, em, dm, xm)
    assert (a[:, :8] - b[:, :8]).abs().max().item() < 1e-6, "decoder must be causal"
    return True


def test_bert_padding_and_t5_causality():
    run_distributed(_bert_t5, 1)


def _mamba(rank, world):
    _init()
    from megatron_b200.core.models.mamba.mamba_layer_specs import mamba_stack_spec
    from megatron_b200.core.models.mamba.mamba_model import MambaModel
    from megatron_b200.core.ssm.ssd import ssd_chunk_scan, ssd_reference
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    torch.manual_seed(0)
    b, l, h, p, g, n = 2, 70, 4, 8, 2, 16
    x, dt = torch.randn(b, l, h, p), torch.nn.functional.softplus(torch.randn(b, l, h))
    A, B, C, D = -torch.rand(h) - 0.5, torch.randn(b, l, g, n), torch.randn(b, l, g, n), torch.randn(h)
    y1, s1 = ssd_chunk_scan(x, dt, A, B, C, 16, D, return_final_states=True)
    y2, s2 = ssd_reference(x, dt, A, B, C, D)
    assert (y1 - y2).abs().max().item() < 1e-3 and (s1 - s2).abs().max().item() < 1e-4
    cfg = TransformerConfig(num_layers=4, hidden_size=64, num_attention_heads=4, normalization="RMSNorm", add_bias_linear=False, **_KW)
    cfg.mamba_state_dim, cfg.mamba_head_dim, cfg.mamba_num_groups = 16, 16, 2
    m = MambaModel(cfg, mamba_stack_spec, 128, 64, hybrid_override_pattern="M*M-", position_embedding_type="rope")
    tok = torch.randint(0, 128, (2, 24))
    pos = torch.arange(24)[None].expand(2, -1)
    m(tok, pos, None, labels=tok).mean().backward()
    assert all(p.grad is not None for p in m.parameters())
    assert _incremental_matches_full(m, tok, pos, 16) < 1e-4  # conv window + SSM state carried across decode steps
    return True


def test_hybrid_mamba_scan_and_decode():
    run_distributed(_mamba, 1)


def _llava(rank, world):
    _init()
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.multimodal.llava_model import LLaVAModel
    from megatron_b200.core.models.vision.vit_layer_specs import get_vit_layer_with_local_spec
    from megatron_b200.core.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
    from megatron_b200.core.transformer.mlp import MLPSubmodules
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    lc = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, **_KW)
    vc = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, **_KW)
    pc = TransformerConfig(num_layers=1, hidden_size=64, num_attention_heads=4, ffn_hidden_size=64, **_KW)
    m = LLaVAModel(lc, get_gpt_layer_local_spec(), 128, 128, vc, get_vit_layer_with_local_spec(), True, pc,
                   MLPSubmodules(ColumnParallelLinear, RowParallelLinear), img_h=28, img_w=28, patch_dim=14)
    tok = torch.randint(1, 128, (2, 12))
    tok[0, 3] = tok[1, 5] = -200
    pos = torch.arange(12)[None].expand(2, -1)
    loss, mask = m(torch.randn(2, 3, 28, 28), tok, pos, None, labels=torch.randint(0, 128, (2, 12)), loss_mask=torch.ones(2, 12))
    assert loss.shape == (2, 15) and mask.sum().item() == 22  # 4 patch embeddings replace each placeholder; images are not predicted
    ((loss * mask).sum() / mask.sum()).backward()
    assert all(p.grad is not None for p in m.parameters())
    return True


def test_llava_splices_image_embeddings():
    run_distributed(_llava, 1)


class _Blk(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b = torch.nn.Linear(16, 32), torch.nn.Linear(32, 16)

    def forward(self, x):
        return x + self.b(torch.tanh(self.a(x)))


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.inp, self.blocks, self.out = torch.nn.Linear(8, 16), torch.nn.ModuleList([_Blk() for _ in range(3)]), torch.nn.Linear(16, 4)

    def forward(self, x):
        x = self.inp(x)
        for b in self.blocks:
            x = b(x)
        return self.out(x)


def _fsdp(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.distributed.fsdp import FullyShardedDataParallel

    torch.manual_seed(0)
    ref, net = _Net(), _Net()
    net.load_state_dict(ref.state_dict())
    f = FullyShardedDataParallel(None, None, net, fsdp_unit_modules=(_Blk,), group=dist.group.WORLD)
    opt, ropt = torch.optim.SGD(f.optimizer_parameters(), lr=0.1), torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(3):
        torch.manual_seed(100 + step)
        X, Y = torch.randn(world * 4, 8), torch.randn(world * 4, 4)
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        ropt.step()
        f.zero_grad_buffer()
        for mb in range(2):
            lo = rank * 4 + mb * 2
            (((f(X[lo : lo + 2]) - Y[lo : lo + 2]) ** 2).mean() / 2).backward()
        f.finish_grad_sync()
        opt.step()
        f.post_optimizer_step()
    assert sum(u.resident for u in f.units) == 1, "only the root unit may stay materialised between steps"
    sd = f.gather_full_state_dict()
    return max((sd[k] - v).abs().max().item() for k, v in ref.state_dict().items())


def test_fsdp_zero3_matches_full_batch_training():
    errs = run_distributed(_fsdp, 2)
    assert max(errs) < 1e-5


def _export_distill_rl(rank, world):
    _init()
    from megatron_b200.core.export.hf_llama import hf_llama_to_megatron, megatron_to_hf_llama
    from megatron_b200.models.presets import build_gpt_model
    from megatron_b200.post_training.distillation import DistillationModel, vocab_parallel_kl
    from megatron_b200.rl.grpo import group_advantages

    model, cfg, p = build_gpt_model("tiny_llama", use_cpu_initialization=True)
    sd = {k: v for k, v in model.state_dict().items() if isinstance(v, torch.Tensor) and "_extra_state" not in k}
    hf = megatron_to_hf_llama(sd, cfg.num_attention_heads, cfg.num_query_groups, cfg.kv_channels)
    assert hf["model.layers.0.self_attn.k_proj.weight"].shape == (cfg.num_query_groups * cfg.kv_channels, cfg.hidden_size)
    back = hf_llama_to_megatron(hf, cfg.num_attention_heads, cfg.num_query_groups, cfg.kv_channels, tie_embeddings="output_layer.weight" not in sd)
    for k, v in back.items():
        assert torch.equal(v, sd[k]), k
    # KD: zero when teacher == student, positive otherwise, gradients reach the student
    a, b = torch.randn(2, 5, 16, requires_grad=True), torch.randn(2, 5, 16)
    assert vocab_parallel_kl(a, a.detach()).abs().max().item() < 1e-6
    kl = vocab_parallel_kl(a, b, temperature=2.0)
    assert (kl > 0).all()
    teacher, _, _ = build_gpt_model("tiny_llama", use_cpu_initialization=True)
    dm = DistillationModel(model, teacher, alpha=0.5)
    tok = torch.randint(0, 100, (2, 16))
    pos = torch.arange(16)[None].expand(2, -1)
    loss, stats = dm(tok, pos, None, tok)
    loss.backward()
    assert all(q.grad is None for q in teacher.parameters()) and any(q.grad is not None for q in model.parameters())
    adv = group_advantages(torch.tensor([1.0, 0.0, 0.5, 0.5, 2.0, 2.0, 2.0, 2.0]), 4)
    assert abs(adv[:4].mean().item()) < 1e-6 and adv[4:].abs().max().item() < 1e-3
    return True


def test_hf_export_roundtrip_distillation_and_grpo_advantages():
    run_distributed(_export_distill_rl, 1)


def _fsdp_ckpt_save(rank, world, path):
    import torch.distributed as dist

    from megatron_b200.core.dist_checkpointing import serialization
    from megatron_b200.core.distributed.fsdp import FullyShardedDataParallel
    from megatron_b200.core.transformer import fsdp_dtensor_checkpoint as fc

    torch.manual_seed(0)
    net = _Net()
    f = FullyShardedDataParallel(None, None, net, fsdp_unit_modules=(_Blk,), group=dist.group.WORLD)
    opt = torch.optim.Adam(f.optimizer_parameters(), lr=0.05)
    for step in range(2):
        torch.manual_seed(100 + step)
        X, Y = torch.randn(world * 2, 8), torch.randn(world * 2, 4)
        f.zero_grad_buffer()
        ((f(X[rank * 2 : rank * 2 + 2]) - Y[rank * 2 : rank * 2 + 2]) ** 2).mean().backward()
        f.finish_grad_sync()
        opt.step()
        f.post_optimizer_step()
    moments = {"exp_avg": [opt.state[u.master]["exp_avg"] for u in f.units]}
    sd = fc.fsdp_model_space_sharded_state_dict(f, extra_states=moments)
    assert all(st.global_shape == (dict(net.named_parameters())[k.split(".exp_avg")[0]].numel(),) for k, st in sd.items())
    serialization.save(sd, path)
    full = f.gather_full_state_dict()
    # same world, different content → load restores
    with torch.no_grad():
        for u in f.units:
            u.master.data.add_(1.0)
    fc.load_fsdp_model_space(f, path)
    assert fc.validate_loaded_state_dict(f, full) == []
    if rank == 0:
        torch.save({k: v for k, v in full.items()}, path + "_full.pt")
    return True


def _fsdp_ckpt_load(rank, world, path, plain):
    import torch.distributed as dist

    full = torch.load(path + "_full.pt")

    from megatron_b200.core.distributed.fsdp import FullyShardedDataParallel
    from megatron_b200.core.transformer import fsdp_dtensor_checkpoint as fc

    torch.manual_seed(9)                                         # different init: everything must come from the checkpoint
    net = _Net()
    if plain:
        fc.load_plain_module_from_model_space(net, path)
        got = net.state_dict()
    else:
        # different world size AND different unit boundaries (whole model as one unit) than at save time
        f = FullyShardedDataParallel(None, None, net, fsdp_unit_modules=(), group=dist.group.WORLD)
        extra = {"exp_avg": [torch.zeros_like(u.master.data) for u in f.units]}
        fc.load_fsdp_model_space(f, path, extra_states=extra)
        got = f.gather_full_state_dict()
        assert float(sum(e.abs().sum() for e in extra["exp_avg"])) > 0
    return max((got[k].float() - v.float()).abs().max().item() for k, v in full.items())


def test_fsdp_model_space_checkpoint_reshards_across_world_sizes_and_into_plain_modules(tmp_path):
    from megatron_b200.core.transformer import fsdp_dtensor_checkpoint as fc

    assert all(run_distributed(_fsdp_ckpt_save, 2, str(tmp_path / "ck")))
    assert max(run_distributed(_fsdp_ckpt_load, 3, str(tmp_path / "ck"), False)) == 0.0
    assert max(run_distributed(_fsdp_ckpt_load, 1, str(tmp_path / "ck"), True)) == 0.0
    assert fc.expert_param_global_key("d.mlp.experts.local_experts.1.linear_fc1.weight", 2, 4) == "d.mlp.experts.9.linear_fc1.weight"
    assert fc.expert_param_local_key("d.mlp.experts.9.linear_fc1.weight", 2, 4) == "d.mlp.experts.local_experts.1.linear_fc1.weight"
    assert fc.expert_param_local_key("d.mlp.experts.9.linear_fc1.weight", 0, 4) is None and fc.get_expert_index_from_key("a.experts.7.w") == 7
    assert fc.flatten_state_dict({"a": {"b": [1, {"c": 2}]}}) == {"a.b.0": 1, "a.b.1.c": 2}
    assert fc.print_diff_in_state_dicts(["a", "b"], ["b", "c"]) == (["c"], ["a"]) and fc._strip_wrapper_prefixes("module.module.x.module.y") == "x.y"
