"""BASELINE config 1 (plumbing): GPT TP=2 (+SP) on CPU/gloo matches the single-process model."""
import torch
import torch.nn.functional as F

from dist_utils import run_distributed


def _build(tp, sp, seed=1234, layers=2, h=64, heads=4, groups=2, ffn=128, vocab=128, seq=32, swiglu=True, norm="RMSNorm", **extra):
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    cfg = TransformerConfig(
        num_layers=layers, hidden_size=h, num_attention_heads=heads, num_query_groups=groups, ffn_hidden_size=ffn,
        use_cpu_initialization=True, normalization=norm, gated_linear_unit=swiglu, activation_func=F.silu if swiglu else F.gelu,
        add_bias_linear=not swiglu, hidden_dropout=0.0, attention_dropout=0.0, tensor_model_parallel_size=tp, sequence_parallel=sp, **extra,
    )
    torch.manual_seed(seed)
    return GPTModel(cfg, get_gpt_layer_local_spec(normalization=norm), vocab_size=vocab, max_sequence_length=seq,
                    position_embedding_type="rope" if swiglu else "learned_absolute"), cfg


def _data(vocab=128, seq=32, b=2):
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, vocab, (b, seq), generator=g)
    labels = torch.randint(0, vocab, (b, seq), generator=g)
    pos = torch.arange(seq)[None].expand(b, -1)
    return ids, labels, pos


def _shard_from_full(name, full, param, rank, world, swiglu):
    """Slice the TP=1 tensor the way a TP rank stores it."""
    if not getattr(param, "tensor_model_parallel", False):
        return full
    dim = param.partition_dim
    if swiglu and "linear_fc1" in name:
        gate, up = full.chunk(2, 0)
        return torch.cat([gate.chunk(world, 0)[rank], up.chunk(world, 0)[rank]], 0)
    return full.chunk(world, dim)[rank]


def _ref_run(swiglu, norm):
    model, _ = _build(1, False, swiglu=swiglu, norm=norm)
    global _REF_STATE
    _REF_STATE = {n: p.detach().clone() for n, p in model.named_parameters()}
    ids, labels, pos = _data()
    loss = model(ids, pos, None, labels=labels).mean()
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    return loss.item(), grads


def _tp_worker(rank, world, sp, swiglu, norm, ref_state):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.distributed.finalize_model_grads import _allreduce_non_tensor_model_parallel_grads
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    model_parallel_cuda_manual_seed(123)
    model, cfg = _build(world, sp, swiglu=swiglu, norm=norm)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(_shard_from_full(n, ref_state[n], p, rank, world, swiglu))
    ids, labels, pos = _data()
    loss = model(ids, pos, None, labels=labels).mean()
    loss.backward()
    _allreduce_non_tensor_model_parallel_grads([model], cfg, ps.get_tensor_model_parallel_group())
    out = {}
    for n, p in model.named_parameters():
        out[n] = (p.grad.clone(), getattr(p, "tensor_model_parallel", False), getattr(p, "partition_dim", -1))
    return loss.item(), out


def _check(sp, swiglu=True, norm="RMSNorm"):
    ref_loss, ref_grads = _ref_run(swiglu, norm)
    res = run_distributed(_tp_worker, 2, sp, swiglu, norm, _REF_STATE)
    for r in range(2):
        assert abs(res[r][0] - ref_loss) < 1e-4, (res[r][0], ref_loss)
    for name, g_ref in ref_grads.items():
        parts = [res[r][1][name] for r in range(2)]
        if parts[0][1]:
            dim = parts[0][2]
            if "linear_fc1.weight" in name and swiglu:
                # per-rank layout [gate_r; up_r] → global [gate_0, gate_1, up_0, up_1]
                gs = [p[0].chunk(2, 0) for p in parts]
                got = torch.cat([gs[0][0], gs[1][0], gs[0][1], gs[1][1]], 0)
            else:
                got = torch.cat([p[0] for p in parts], dim)
        else:
            got = parts[0][0]
            assert torch.allclose(parts[0][0], parts[1][0], atol=1e-5), name
        assert torch.allclose(got, g_ref, atol=2e-4, rtol=1e-3), (name, (got - g_ref).abs().max())


def test_gpt_tp2_matches_single():
    _check(sp=False)


def test_gpt_tp2_sequence_parallel_matches_single():
    _check(sp=True)


def test_gpt2_style_tp2_sp_layernorm_gelu():
    _check(sp=True, swiglu=False, norm="LayerNorm")


def _selfcheck_worker(rank, world):
    from megatron_b200.parallel.selfcheck import pair_op_self_check
    import torch.distributed as dist

    return pair_op_self_check(dist.group.WORLD, seq=world * 64, hidden=64, ffn=128, qkv=96, quick=False, repeats=2)


def test_pair_op_self_check_gloo():
    """The bench's fused-vs-collective self check runs (trivially passes) on the gloo decomposition of the pair ops."""
    from dist_utils import run_distributed

    res = run_distributed(_selfcheck_worker, 2)
    assert res[0]["max"] < 1e-4 and len(res[0]) >= 12, res[0]
