"""Pure-function and single-process CPU tests: rank grids, datasets/native helpers, schedulers,
optimizer math, checkpoint mapping, FLOPs model."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch


# ---------------------------------------------------------------------------------- rank generator
def test_rank_generator_default_order():
    from megatron_b200.core.parallel_state import RankGenerator

    g = RankGenerator(tp=2, ep=1, dp=2, pp=2, cp=1, order="tp-cp-ep-dp-pp")
    assert g.get_ranks("tp") == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert g.get_ranks("dp") == [[0, 2], [1, 3], [4, 6], [5, 7]]
    assert g.get_ranks("pp") == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert g.get_ranks("tp-pp") == [[0, 1, 4, 5], [2, 3, 6, 7]]
    assert g.get_ranks("tp-dp") == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_rank_generator_cp_and_expert():
    from megatron_b200.core.parallel_state import RankGenerator

    g = RankGenerator(tp=2, ep=1, dp=1, pp=1, cp=4, order="tp-cp-ep-dp-pp")
    assert g.get_ranks("cp") == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert g.get_ranks("dp-cp") == [[0, 2, 4, 6], [1, 3, 5, 7]]
    e = RankGenerator(tp=1, ep=4, dp=2, pp=1, cp=1, order="tp-cp-ep-dp-pp")
    assert e.get_ranks("ep") == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert e.get_ranks("dp") == [[0, 4], [1, 5], [2, 6], [3, 7]]
    with pytest.raises(RuntimeError):
        RankGenerator(tp=2, ep=1, dp=2, pp=2, cp=1, order="tp-dp")  # pp=2 missing from order


def test_rank_groups_partition_the_world():
    from megatron_b200.core.parallel_state import RankGenerator

    g = RankGenerator(tp=2, ep=1, dp=3, pp=2, cp=2, order="tp-cp-ep-dp-pp")
    for token in ("tp", "dp", "pp", "cp", "tp-cp", "dp-cp", "tp-dp-cp", "tp-pp"):
        flat = sorted(r for grp in g.get_ranks(token) for r in grp)
        assert flat == list(range(24)), token


# ---------------------------------------------------------------------------------- datasets
def _py_sample_idx(sizes, doc_idx, seq, num_epochs, tokens_per_epoch, extra=1):
    """Reference implementation by brute force: materialise the token→(doc, offset) stream."""
    stream = []
    for i, d in enumerate(doc_idx):
        stream += [(i, o) for o in range(sizes[d])]
    n = (num_epochs * tokens_per_epoch - extra) // seq
    out = [stream[s * seq] for s in range(n)]
    last = n * seq
    out.append(stream[last] if last < len(stream) else (len(doc_idx) - 1, sizes[doc_idx[-1]] - extra))
    return np.array(out)


def test_native_build_sample_idx_matches_bruteforce():
    from megatron_b200.core.datasets import helpers

    rng = np.random.RandomState(0)
    for trial in range(20):
        ndoc = rng.randint(3, 40)
        sizes = rng.randint(1, 50, size=ndoc).astype(np.int32)
        epochs = rng.randint(1, 4)
        doc_idx = np.tile(np.arange(ndoc, dtype=np.int32), epochs)
        rng.shuffle(doc_idx)
        seq = int(rng.randint(2, 30))
        tpe = int(sizes.sum())
        if (epochs * tpe - 1) // seq < 1:
            continue
        got = helpers.build_sample_idx(sizes, doc_idx, seq, epochs, tpe, True, True)
        ref = _py_sample_idx(sizes, doc_idx, seq, epochs, tpe)
        assert got.shape == ref.shape, (got.shape, ref.shape)
        # boundaries may be expressed as (doc i, len_i) or (doc i+1, 0): compare absolute token positions
        starts = np.concatenate([[0], np.cumsum(sizes[doc_idx])])
        pos = lambda a: starts[a[:, 0]] + a[:, 1]
        assert np.array_equal(pos(got)[:-1], pos(ref)[:-1]), trial


def test_blending_indices_follow_weights():
    from megatron_b200.core.datasets import helpers

    n, w = 10000, [0.5, 0.3, 0.2]
    di, dsi = np.zeros(n, dtype=np.int16), np.zeros(n, dtype=np.int64)
    helpers.build_blending_indices(di, dsi, w, 3, n, False)
    for d in range(3):
        assert abs((di == d).mean() - w[d]) < 2e-3
        assert np.array_equal(dsi[di == d], np.arange((di == d).sum()))
    sizes = [7, 3, 5]
    di, dsi = np.zeros(15, dtype=np.int16), np.zeros(15, dtype=np.int64)
    helpers.build_exhaustive_blending_indices(di, dsi, sizes, 3)
    assert [int((di == d).sum()) for d in range(3)] == sizes


def test_indexed_dataset_roundtrip_and_gpt_dataset():
    from megatron_b200.core.datasets import GPTDatasetConfig, IndexedDataset, IndexedDatasetBuilder
    from megatron_b200.core.datasets.blended_megatron_dataset_builder import BlendedMegatronDatasetBuilder
    from megatron_b200.core.datasets.gpt_dataset import GPTDataset
    from megatron_b200.core.tokenizers import NullTokenizer

    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "corpus")
        b = IndexedDatasetBuilder(prefix + ".bin", dtype=np.uint16)
        rng = np.random.RandomState(1)
        docs = [rng.randint(1, 200, size=rng.randint(5, 60)) for _ in range(200)]
        for doc in docs:
            b.add_item(torch.tensor(np.append(doc, 255)))
            b.end_document()
        b.finalize(prefix + ".idx")
        ds = IndexedDataset(prefix)
        assert len(ds) == 200 and np.array_equal(ds[3][:-1], docs[3]) and np.array_equal(ds.get(3, 2, 3), docs[3][2:5])
        cfg = GPTDatasetConfig(random_seed=7, sequence_length=32, blend=([prefix], None), split="90,5,5", reset_position_ids=False,
                               reset_attention_mask=False, eod_mask_loss=True, tokenizer=NullTokenizer(255), path_to_cache=os.path.join(d, "cache"))
        train, valid, test = BlendedMegatronDatasetBuilder(GPTDataset, [300, 10, 10], lambda: True, cfg).build()
        assert len(train) >= 300
        s = train[0]
        assert s["tokens"].shape == (32,) and torch.equal(s["tokens"][1:], s["labels"][:-1])
        assert (s["loss_mask"][s["tokens"] == 255] == 0).all()
        again, _, _ = BlendedMegatronDatasetBuilder(GPTDataset, [300, 10, 10], lambda: True, cfg).build()  # from cache
        assert torch.equal(again[17]["tokens"], train[17]["tokens"])


# ---------------------------------------------------------------------------------- schedulers / calculators
def test_lr_scheduler_shapes():
    from megatron_b200.core.optimizer_param_scheduler import OptimizerParamScheduler

    class Opt:
        param_groups = [{"lr": 0.0, "weight_decay": 0.0}]

    o = Opt()
    s = OptimizerParamScheduler(o, init_lr=0.0, max_lr=1.0, min_lr=0.1, lr_warmup_steps=10, lr_decay_steps=110, lr_decay_style="cosine",
                                start_wd=0.1, end_wd=0.1, wd_incr_steps=110, wd_incr_style="constant", use_checkpoint_opt_param_scheduler=False)
    s.step(5)
    assert abs(o.param_groups[0]["lr"] - 0.5) < 1e-9
    s.step(5)
    assert abs(o.param_groups[0]["lr"] - 1.0) < 1e-9
    s.step(50)
    assert abs(o.param_groups[0]["lr"] - (0.1 + 0.9 * 0.5 * (math.cos(math.pi * 0.5) + 1))) < 1e-9
    s.step(1000)
    assert o.param_groups[0]["lr"] == 0.1
    sd = s.state_dict()
    s2 = OptimizerParamScheduler(Opt(), 0.0, 1.0, 0.1, 10, 110, "cosine", 0.1, 0.1, 110, "constant", use_checkpoint_opt_param_scheduler=True)
    s2.load_state_dict(sd)
    assert s2.num_steps == s.num_steps


def test_num_microbatches_rampup():
    from megatron_b200.core import num_microbatches_calculator as nm

    nm.destroy_num_microbatches_calculator()
    nm.init_num_microbatches_calculator(0, [8, 8, 64], 32, 2, 2)
    assert nm.get_num_microbatches() == 2 and nm.get_current_global_batch_size() == 8
    nm.update_num_microbatches(40)
    assert nm.get_current_global_batch_size() == 16 + 8 * 0 or nm.get_current_global_batch_size() in (16, 24)
    nm.update_num_microbatches(1000)
    assert nm.get_current_global_batch_size() == 32 and nm.get_num_microbatches() == 8
    nm.destroy_num_microbatches_calculator()


def test_flops_model_matches_reference_accounting():
    from megatron_b200.models.presets import PRESETS
    from megatron_b200.training.flops import flops_per_token

    p = PRESETS["llama3_8b"]
    f = flops_per_token(**{k: p[k] for k in ("num_layers", "hidden_size", "ffn_hidden_size", "num_attention_heads", "num_query_groups", "kv_channels", "vocab_size", "seq_length")})
    assert abs(f / 1e9 - 51.5) < 0.1  # ≈ 51.5 GFLOP per token


# ---------------------------------------------------------------------------------- ops references / optimizer
def test_fused_adam_reference_matches_torch_adamw():
    from megatron_b200 import ops

    torch.manual_seed(0)
    p = torch.randn(1000)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    m, v, mine = torch.zeros(1000), torch.zeros(1000), p.clone()
    for step in range(1, 6):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        ops.fused_adam([mine], [g], [m], [v], [None], lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=step, adamw=True)
    assert torch.allclose(mine, ref.detach(), atol=1e-6)


def test_rms_norm_and_swiglu_and_rope_autograd_cpu():
    from megatron_b200 import ops

    torch.manual_seed(0)
    def same_grads(fn_ours, fn_eager, inputs, tol=1e-4):
        """closed-form backward of the op == autograd through an eager formulation"""
        a = [t.clone().requires_grad_() for t in inputs]
        b = [t.clone().requires_grad_() for t in inputs]
        ya, yb = fn_ours(*a), fn_eager(*b)
        assert torch.allclose(ya, yb, atol=tol), (ya - yb).abs().max()
        g = torch.randn_like(ya)
        for ga, gb in zip(torch.autograd.grad(ya, a, g), torch.autograd.grad(yb, b, g)):
            assert torch.allclose(ga, gb, atol=tol), (ga - gb).abs().max()

    x, w = torch.randn(6, 32), torch.randn(32)
    same_grads(lambda a, b: ops.rms_norm(a, b, 1e-5), lambda a, b: a * torch.rsqrt(a.pow(2).mean(-1, keepdim=True) + 1e-5) * b, (x, w))
    lb = torch.randn(32)
    same_grads(lambda a, b, c: ops.layer_norm(a, b, c, 1e-5), lambda a, b, c: torch.nn.functional.layer_norm(a, (32,), b, c, 1e-5), (x, w, lb))
    y, pr = torch.randn(5, 16), torch.rand(5, 1)
    same_grads(lambda a, b: ops.swiglu(a, None, b), lambda a, b: torch.nn.functional.silu(a[:, :8]) * a[:, 8:] * b, (y, pr))
    half = torch.randn(4, 1, 1, 4)
    t, fr = torch.randn(4, 2, 3, 8), torch.cat((half, half), -1)  # RoPE angles repeat across the two halves

    def rope_eager(a):
        x1, x2 = a.chunk(2, -1)
        return a * torch.cos(fr) + torch.cat((-x2, x1), -1) * torch.sin(fr)

    same_grads(lambda a: ops.apply_rope(a, fr), rope_eager, (t,))


def test_vocab_parallel_ce_matches_torch_cpu():
    from megatron_b200 import ops

    torch.manual_seed(0)
    logits = torch.randn(4, 3, 50, requires_grad=True)
    target = torch.randint(0, 50, (4, 3))
    loss = ops.vocab_parallel_cross_entropy(logits, target)
    ref = torch.nn.functional.cross_entropy(logits.reshape(-1, 50), target.reshape(-1), reduction="none").view(4, 3)
    assert torch.allclose(loss, ref, atol=1e-5)
    g = torch.autograd.grad(loss.sum(), logits)[0]
    gr = torch.autograd.grad(ref.sum(), logits)[0]
    assert torch.allclose(g, gr, atol=1e-5)
    ls = ops.vocab_parallel_cross_entropy(logits, target, label_smoothing=0.1)
    refs = torch.nn.functional.cross_entropy(logits.reshape(-1, 50), target.reshape(-1), reduction="none", label_smoothing=0.1 * 50 / 49).view(4, 3)
    assert torch.allclose(ls, refs, atol=1e-4)


# ---------------------------------------------------------------------------------- checkpoint mapping
def test_sharded_tensor_from_rank_offsets_and_swiglu_factory():
    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor, apply_factories
    from megatron_b200.core.transformer.mlp import apply_swiglu_sharded_factory

    t = torch.arange(24.0).view(4, 6)
    st = ShardedTensor.from_rank_offsets("w", t, (0, 1, 2), (1, 4, 8), prepend_axis_num=1)
    # axis 0 is the prepended (layer) axis: 2 layers, this is layer 1; axis 1 = first data axis, piece 4 of 8
    assert st.global_shape == (2, 32, 6) and st.global_offset == (1, 16, 0) and st.local_shape == (4, 6)
    st2 = ShardedTensor.from_rank_offsets("fc1", t, (0, 1, 4))  # tp rank 1 of 4
    fac = apply_swiglu_sharded_factory(st2, ())
    sd = {"x": fac}
    apply_factories(sd)
    gate, up = sd["x"]
    assert gate.global_shape == (16, 6) and gate.global_offset == (2, 0) and up.global_offset == (10, 0)
    assert torch.equal(fac.merge_fn([gate.data, up.data]), t)
    with pytest.raises(ValueError):
        ShardedTensor.from_rank_offsets("bad", t, (0, 2, 2))


def test_checkpoint_detects_double_writer_and_holes():
    from megatron_b200.core.dist_checkpointing.core import CheckpointingException
    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor
    from megatron_b200.core.dist_checkpointing.validation import validate_sharding_integrity

    t = torch.zeros(2, 2)
    ok = {"a": ShardedTensor.from_rank_offsets("k", t, (0, 0, 2)), "b": ShardedTensor.from_rank_offsets("k", t, (0, 1, 2))}
    validate_sharding_integrity(ok)
    with pytest.raises(CheckpointingException):
        validate_sharding_integrity({"a": ShardedTensor.from_rank_offsets("k", t, (0, 0, 2))})  # hole
    with pytest.raises(CheckpointingException):
        validate_sharding_integrity({"a": ShardedTensor.from_rank_offsets("k", t), "b": ShardedTensor.from_rank_offsets("k", t)})  # two writers


def test_fully_parallel_save_assignment_is_balanced():
    from megatron_b200.core.dist_checkpointing.strategies.fully_parallel import distribute_shards_to_ranks

    shards = {f"s{i}": [0, 1, 2, 3] for i in range(8)}
    sizes = {f"s{i}": 100 for i in range(8)}
    a = distribute_shards_to_ranks(shards, sizes, 4)
    loads = [sum(sizes[s] for s, r in a.items() if r == k) for k in range(4)]
    assert loads == [200] * 4
    a = distribute_shards_to_ranks({"big": [0, 1], "only1": [1]}, {"big": 10, "only1": 10}, 2)
    assert a["only1"] == 1 and a["big"] == 0


def test_moe_permute_unpermute_roundtrip():
    from megatron_b200.core.transformer.moe.moe_utils import permute, topk_routing_with_score_function, unpermute

    torch.manual_seed(0)
    T, E, H = 17, 4, 8
    x = torch.randn(T, H)
    probs, rmap = topk_routing_with_score_function(torch.randn(T, E), 2)
    px, pp, idx = permute(x, rmap, probs)
    assert px.shape[0] == 2 * T and torch.allclose(pp.sum(), probs.sum())
    back = unpermute(px, idx, x.shape, probs=probs, routing_map=rmap)
    assert torch.allclose(back, x * probs.sum(-1, keepdim=True), atol=1e-5)


def test_inference_kv_cache_matches_full_forward():
    from megatron_b200.core.inference_params import InferenceParams
    from megatron_b200.models.presets import build_gpt_model

    torch.manual_seed(0)
    m, cfg, p = build_gpt_model("tiny_llama", use_cpu_initialization=True, num_layers=2)
    m.eval()
    ids = torch.randint(0, 1024, (1, 12))
    pos = torch.arange(12)[None]
    with torch.no_grad():
        full = m(ids, pos, None)
        ctx = InferenceParams(1, 32)
        out = m(ids[:, :8], pos[:, :8], None, inference_context=ctx)
        ctx.sequence_len_offset = 8
        outs = [out]
        for t in range(8, 12):
            o = m(ids[:, t : t + 1], pos[:, t : t + 1], None, inference_context=ctx)
            ctx.sequence_len_offset += 1
            outs.append(o)
        inc = torch.cat(outs, dim=1)
    assert torch.allclose(full, inc, atol=1e-4), (full - inc).abs().max()


def test_remove_sharded_tensors(tmp_path):
    import torch

    from megatron_b200.core import dist_checkpointing as dc
    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor
    from megatron_b200.core.dist_checkpointing.serialization import load_tensors_metadata, remove_sharded_tensors

    d = tmp_path / "ck"
    d.mkdir()
    sd = {"a.w": ShardedTensor.from_rank_offsets("a.w", torch.ones(4, 4)), "opt.m": ShardedTensor.from_rank_offsets("opt.m", torch.zeros(4)), "opt.v": ShardedTensor.from_rank_offsets("opt.v", torch.zeros(4))}
    dc.save(sd, str(d))
    assert set(load_tensors_metadata(str(d))) == {"a.w", "opt.m", "opt.v"}
    remove_sharded_tensors(str(d), "opt.")
    assert set(load_tensors_metadata(str(d))) == {"a.w"}
    out = dc.load({"a.w": ShardedTensor.from_rank_offsets("a.w", torch.zeros(4, 4))}, str(d))
    assert torch.equal(out["a.w"], torch.ones(4, 4))
