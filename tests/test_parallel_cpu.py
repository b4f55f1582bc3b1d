"""Multi-process (gloo) equivalence tests: pipeline schedules, expert parallel, data parallel +
distributed optimizer, checkpoint save → reshard → load."""
import os
import tempfile

import pytest
import torch
import torch.nn.functional as F

from dist_utils import run_distributed

SEQ, VOCAB = 32, 128


def _cfg(**kw):
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    base = dict(num_layers=4, hidden_size=64, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=128, use_cpu_initialization=True,
                normalization="RMSNorm", gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False, hidden_dropout=0.0, attention_dropout=0.0)
    base.update(kw)
    return TransformerConfig(**base)


def _model(cfg, pre=True, post=True, vp_stage=None, tie=False):
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_decoder_block_spec, get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel

    spec = get_gpt_decoder_block_spec(cfg, vp_stage=vp_stage) if cfg.num_moe_experts else get_gpt_layer_local_spec(normalization="RMSNorm")
    return GPTModel(cfg, spec, vocab_size=VOCAB, max_sequence_length=SEQ, pre_process=pre, post_process=post, position_embedding_type="rope",
                    share_embeddings_and_output_weights=tie, vp_stage=vp_stage)


def _batches(n, b=2, seed=3):
    g = torch.Generator().manual_seed(seed)
    return [dict(tokens=torch.randint(0, VOCAB, (b, SEQ), generator=g), labels=torch.randint(0, VOCAB, (b, SEQ), generator=g)) for _ in range(n)]


def _fstep(it, model):
    b = next(it)
    pos = torch.arange(SEQ)[None].expand(b["tokens"].shape[0], -1)
    out = model(b["tokens"], pos, None, labels=b["labels"])

    def loss_fn(o):
        l = o.float().mean()
        return l, {"lm loss": l.detach()}

    return out, loss_fn


def _reference(num_mb, seed=11, **cfgkw):
    torch.manual_seed(seed)
    m = _model(_cfg(**cfgkw))
    state = {n: p.detach().clone() for n, p in m.named_parameters()}
    losses = []
    for b in _batches(num_mb):
        out, lf = _fstep(iter([b]), m)
        l, _ = lf(out)
        (l / num_mb).backward()
        losses.append(l.item())
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    return state, losses, grads


# ------------------------------------------------------------------------------- pipeline
def _pp_worker(rank, world, vp, num_mb, ref_state):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.pipeline_parallel.schedules import get_forward_backward_func
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_layer import get_transformer_layer_offset

    ps.initialize_model_parallel(pipeline_model_parallel_size=world, virtual_pipeline_model_parallel_size=vp)
    model_parallel_cuda_manual_seed(1)
    cfg = _cfg(pipeline_model_parallel_size=world, virtual_pipeline_model_parallel_size=vp, pipeline_dtype=torch.float32)
    chunks = []
    for v in range(vp or 1):
        if vp:
            ps.set_virtual_pipeline_model_parallel_rank(v)
        pre = ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=v if vp else None)
        post = ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=v if vp else None)
        m = _model(cfg, pre, post, vp_stage=v if vp else None)
        off = get_transformer_layer_offset(cfg, v if vp else None)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.startswith("decoder.layers."):
                    parts = n.split(".")
                    parts[2] = str(int(parts[2]) + off)
                    p.copy_(ref_state[".".join(parts)])
                else:
                    p.copy_(ref_state[n])
        m.config = cfg
        chunks.append((m, off))
    if vp:
        ps.set_virtual_pipeline_model_parallel_rank(0)
    data = _batches(num_mb)
    fb = get_forward_backward_func()
    models = [c[0] for c in chunks]
    its = [iter(data) for _ in models]
    losses = fb(forward_step_func=_fstep, data_iterator=its if vp else its[0], model=models if vp else models[0], num_microbatches=num_mb,
                seq_length=SEQ, micro_batch_size=2, forward_only=False)
    grads = {}
    for m, off in chunks:
        for n, p in m.named_parameters():
            if n.startswith("decoder.layers."):
                parts = n.split(".")
                parts[2] = str(int(parts[2]) + off)
                n = ".".join(parts)
            grads[n] = p.grad.clone() if p.grad is not None else None
    return [float(d["lm loss"]) for d in losses], grads


@pytest.mark.parametrize("world,vp,num_mb", [(2, None, 4), (4, None, 6), (2, 2, 4), (2, 2, 6), (4, None, 2)])
def test_pipeline_schedules_match_single_process(world, vp, num_mb):
    state, ref_losses, ref_grads = _reference(num_mb)
    res = run_distributed(_pp_worker, world, vp, num_mb, state)
    last = res[world - 1]
    assert len(last[0]) == num_mb
    for a, b in zip(sorted(last[0]), sorted(ref_losses)):
        assert abs(a - b) < 1e-4
    seen = set()
    for r in range(world):
        for n, g in res[r][1].items():
            assert g is not None, f"rank {r} param {n} got no grad"
            assert torch.allclose(g, ref_grads[n], atol=2e-5, rtol=1e-4), (r, n, (g - ref_grads[n]).abs().max())
            seen.add(n)
    assert seen == set(ref_grads)


def test_interleaved_plan_is_well_formed():
    from megatron_b200.core.pipeline_parallel.schedules import build_interleaved_plan, get_schedule_table

    assert get_schedule_table(5, 2, 3) == [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1), (3, 0), (4, 0), (3, 1), (4, 1)]
    for pp, vp, m in [(2, 2, 4), (4, 2, 8), (4, 3, 8), (2, 2, 2)]:
        for rank in range(pp):
            plan = build_interleaved_plan(m, vp, pp, rank, pp)
            f = [(mb, ch) for op, _, mb, ch in plan if op == "F"]
            b = [(mb, ch) for op, _, mb, ch in plan if op == "B"]
            assert sorted(f) == sorted(b) == sorted((i, c) for i in range(m) for c in range(vp))
            done = set()
            for op, _, mb, ch in plan:  # a backward never precedes its forward
                if op == "F":
                    done.add((mb, ch))
                else:
                    assert (mb, ch) in done


# ------------------------------------------------------------------------------- expert parallel
def _ep_worker(rank, world, dispatcher, grouped, ref_state):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(expert_model_parallel_size=world)
    model_parallel_cuda_manual_seed(1)
    cfg = _cfg(num_layers=2, num_moe_experts=4, moe_router_topk=2, moe_token_dispatcher_type=dispatcher, moe_grouped_gemm=grouped,
               expert_model_parallel_size=world, moe_aux_loss_coeff=0.0)
    m = _model(cfg)
    L = 4 // world
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".experts.weight" in n:  # grouped: [L, ...] slice of the [E, ...] reference
                p.copy_(ref_state[n][rank * L : (rank + 1) * L])
            elif ".local_experts." in n:
                parts = n.split(".")
                i = parts.index("local_experts")
                parts[i + 1] = str(int(parts[i + 1]) + rank * L)
                p.copy_(ref_state[".".join(parts)])
            else:
                p.copy_(ref_state[n])
    # every EP rank is also a DP rank: feed different data, compare against the matching reference run
    b = _batches(world, seed=5)[rank]
    out, lf = _fstep(iter([b]), m)
    l, _ = lf(out)
    l.backward()
    return l.item(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("dispatcher,grouped", [("alltoall", True), ("allgather", False), ("alltoall", False), ("flex", True)])
def test_expert_parallel_matches_single_process(dispatcher, grouped):
    world = 2
    torch.manual_seed(21)
    cfg = _cfg(num_layers=2, num_moe_experts=4, moe_router_topk=2, moe_token_dispatcher_type="alltoall", moe_grouped_gemm=grouped, moe_aux_loss_coeff=0.0)
    ref = _model(cfg)
    state = {n: p.detach().clone() for n, p in ref.named_parameters()}
    ref_losses, ref_grads = [], None
    for b in _batches(world, seed=5):
        out, lf = _fstep(iter([b]), ref)
        l, _ = lf(out)
        l.backward()
        ref_losses.append(l.item())
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
    res = run_distributed(_ep_worker, world, dispatcher, grouped, state)
    for r in range(world):
        assert abs(res[r][0] - ref_losses[r]) < 1e-4, (res[r][0], ref_losses[r])
    # expert grads: each rank owns distinct experts and has seen the tokens of BOTH ranks → equals the summed reference grad
    L = 4 // world
    for n, g_ref in ref_grads.items():
        if ".experts.weight" in n:
            got = torch.cat([res[r][1][n] for r in range(world)], 0)
            assert torch.allclose(got, g_ref, atol=2e-5, rtol=1e-4), n
        elif ".local_experts." in n:
            parts = n.split(".")
            i = parts.index("local_experts")
            e = int(parts[i + 1])
            parts[i + 1] = str(e % L)
            got = res[e // L][1][".".join(parts)]
            assert torch.allclose(got, g_ref, atol=2e-5, rtol=1e-4), n
        else:
            got = sum(res[r][1][n] for r in range(world))
            assert torch.allclose(got, g_ref, atol=2e-5, rtol=1e-4), n


# ------------------------------------------------------------------------------- DP + distributed optimizer
def _dp_worker(rank, world, dist_opt, overlap, steps):
    from megatron_b200.training.engine import TrainEngine

    eng = TrainEngine("tiny_llama", micro_batch_size=1, global_batch_size=4, bf16=False, use_distributed_optimizer=dist_opt,
                      overlap_grad_reduce=overlap, overlap_param_gather=False, seed=99, ddp_bucket_size=20000,
                      model_overrides=dict(num_layers=2, hidden_size=64, ffn_hidden_size=128, num_attention_heads=4, num_query_groups=2, kv_channels=16,
                                           vocab_size=256, seq_length=32))
    g = torch.Generator().manual_seed(123)
    full = torch.randint(0, 256, (4, 33), generator=g)
    per = 4 // world
    mine = full[rank * per : (rank + 1) * per]
    losses = [float(eng.train_step(mine)) for _ in range(steps)]
    params = torch.cat([p.detach().reshape(-1) for p in eng.model_chunks[0].parameters()])
    return losses, params


@pytest.mark.parametrize("dist_opt,overlap", [(True, True), (True, False), (False, False)])
def test_data_parallel_and_distributed_optimizer_match_single_process(dist_opt, overlap):
    ref = run_distributed(_dp_worker, 1, dist_opt, False, 3)[0]
    res = run_distributed(_dp_worker, 2, dist_opt, overlap, 3)
    assert torch.allclose(res[0][1], res[1][1], atol=1e-7), "DP replicas diverged"
    assert torch.allclose(res[0][1], ref[1], atol=1e-5, rtol=1e-4), (res[0][1] - ref[1]).abs().max()
    mean_losses = [(a + b) / 2 for a, b in zip(res[0][0], res[1][0])]
    for a, b in zip(mean_losses, ref[0]):
        assert abs(a - b) < 1e-4


# ------------------------------------------------------------------------------- checkpoint resharding
def _ckpt_save_worker(rank, world, tp, pp, path, ref_state):
    from megatron_b200.core import dist_checkpointing as dc
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_layer import get_transformer_layer_offset
    from test_tp_cpu import _shard_from_full

    ps.initialize_model_parallel(tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp)
    model_parallel_cuda_manual_seed(1)
    cfg = _cfg(tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp, pipeline_dtype=torch.float32)
    m = _model(cfg, ps.is_pipeline_first_stage(), ps.is_pipeline_last_stage())
    off = get_transformer_layer_offset(cfg)
    tpr = ps.get_tensor_model_parallel_rank()
    if ref_state is not None:
        with torch.no_grad():
            for n, p in m.named_parameters():
                key = n
                if n.startswith("decoder.layers."):
                    parts = n.split(".")
                    parts[2] = str(int(parts[2]) + off)
                    key = ".".join(parts)
                p.copy_(_shard_from_full(key, ref_state[key], p, tpr, tp, True))
        dc.save(m.sharded_state_dict(), path)
        return None
    sd = dc.load(m.sharded_state_dict(), path)
    m.load_state_dict(sd)
    out = {}
    for n, p in m.named_parameters():
        key = n
        if n.startswith("decoder.layers."):
            parts = n.split(".")
            parts[2] = str(int(parts[2]) + off)
            key = ".".join(parts)
        out[key] = (p.detach().clone(), getattr(p, "tensor_model_parallel", False), getattr(p, "partition_dim", -1), tpr)
    return out


def test_checkpoint_save_tp2_pp2_load_tp1_pp4_and_tp4():
    from test_tp_cpu import _shard_from_full

    torch.manual_seed(5)
    ref = _model(_cfg())
    state = {n: p.detach().clone() for n, p in ref.named_parameters()}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt")
        run_distributed(_ckpt_save_worker, 4, 2, 2, path, state)
        assert os.path.exists(os.path.join(path, ".metadata")) and os.path.exists(os.path.join(path, "metadata.json"))
        for tp, pp in [(1, 4), (4, 1), (2, 1)]:
            res = run_distributed(_ckpt_save_worker, tp * pp, tp, pp, path, None)
            seen = set()
            for r, out in enumerate(res):
                for key, (val, is_tp, dim, tpr) in out.items():
                    class P:  # minimal stand-in carrying the TP attributes for the slicer
                        tensor_model_parallel, partition_dim = is_tp, dim
                    exp = _shard_from_full(key, state[key], P, tpr, tp, True)
                    assert torch.equal(val, exp), (tp, pp, r, key)
                    seen.add(key)
            assert seen == set(state)


# ------------------------------------------------------------------------------- context parallel
def _cp_worker(rank, world, kind, state):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.utils import get_batch_on_this_cp_rank

    ps.initialize_model_parallel(context_parallel_size=world, hierarchical_context_parallel_sizes=[2, world // 2] if kind == "a2a+p2p" else None)
    model_parallel_cuda_manual_seed(1)
    cfg = _cfg(num_layers=2, context_parallel_size=world, cp_comm_type=kind)
    m = _model(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(state[n])
    b = _batches(1, seed=9)[0]
    pos = torch.arange(SEQ)[None].expand(2, -1)
    local = get_batch_on_this_cp_rank({"tokens": b["tokens"], "labels": b["labels"], "position_ids": pos})
    out = m(local["tokens"], local["position_ids"], None, labels=local["labels"])
    loss_sum = out.float().sum()
    (loss_sum / (2 * SEQ)).backward()
    return float(loss_sum), {n: p.grad.clone() for n, p in m.named_parameters()}


@pytest.mark.parametrize("kind", ["all_gather", "p2p", "a2a", "a2a+p2p"])
def test_context_parallel_matches_single_process(kind):
    world = 4 if kind == "a2a+p2p" else 2
    torch.manual_seed(33)
    ref = _model(_cfg(num_layers=2))
    state = {n: p.detach().clone() for n, p in ref.named_parameters()}
    b = _batches(1, seed=9)[0]
    pos = torch.arange(SEQ)[None].expand(2, -1)
    out = ref(b["tokens"], pos, None, labels=b["labels"])
    out.float().mean().backward()
    res = run_distributed(_cp_worker, world, kind, state)
    total = sum(r[0] for r in res) / (2 * SEQ)
    assert abs(total - float(out.float().mean())) < 1e-4
    for n, p in ref.named_parameters():
        got = sum(r[1][n] for r in res)
        assert torch.allclose(got, p.grad, atol=3e-5, rtol=1e-4), (n, (got - p.grad).abs().max())


def _async_process_save_worker(rank, world, ckpt_dir):
    """Async save through the persistent writer PROCESS: the model keeps training (weights are overwritten) while the files are written; the checkpoint
    must hold the values at save time; loading it back at a different TP layout (world 2 -> plain tensors) must reproduce them."""
    import torch

    from megatron_b200.core import dist_checkpointing as dc
    from megatron_b200.core.dist_checkpointing.mapping import ShardedObject, ShardedTensor
    from megatron_b200.core.dist_checkpointing.strategies.async_utils import AsyncCallsQueue, ProcessAsyncRequest

    torch.manual_seed(0)
    full = torch.arange(64 * 8, dtype=torch.float32).reshape(64, 8)
    mine = full.chunk(world, dim=0)[rank].clone()
    repl = torch.full((5,), 3.0)
    sd = {"w": ShardedTensor.from_rank_offsets("w", mine, (0, rank, world)), "r": ShardedTensor.from_rank_offsets("r", repl, replica_id=rank),
          "obj": ShardedObject("obj", {"step": 7}, (1,), (0,), replica_id=rank), "iteration": 7}
    req = dc.save(sd, ckpt_dir, async_sharded_save=True, async_strategy="process")
    assert isinstance(req, ProcessAsyncRequest)
    q = AsyncCallsQueue()
    q.schedule_async_request(req)
    mine.fill_(-1.0)          # training goes on: the staged copy must not alias the live tensor
    repl.fill_(-1.0)
    done = []
    import time

    t0 = time.time()
    while not done and time.time() - t0 < 120:
        done = q.maybe_finalize_async_calls(blocking=False)
        time.sleep(0.05)
    assert done == [1], "async save did not finalize"
    out = {"w": ShardedTensor.from_rank_offsets("w", torch.zeros(64 // world, 8), (0, rank, world)), "r": ShardedTensor.from_rank_offsets("r", torch.zeros(5), replica_id=rank),
           "obj": ShardedObject("obj", None, (1,), (0,), replica_id=rank)}
    loaded = dc.load(out, ckpt_dir)
    assert torch.equal(loaded["w"], full.chunk(world, dim=0)[rank]) and torch.equal(loaded["r"], torch.full((5,), 3.0))
    assert loaded["obj"] == {"step": 7} and loaded["iteration"] == 7
    from megatron_b200.core.dist_checkpointing.strategies.async_utils import PersistentWriterProcess

    PersistentWriterProcess.get().close()
    return True


def test_async_save_in_worker_process(tmp_path):
    from dist_utils import run_distributed

    d = tmp_path / "ck"
    d.mkdir()
    assert all(run_distributed(_async_process_save_worker, 2, str(d)))


def _fully_parallel_load_worker(rank, world, ckpt_dir):
    """4 ranks = TP 2 x DP 2: every TP shard is requested by both DP replicas; with the fully-parallel load strategy each shard is read from storage by ONE
    of them (about half the bytes per rank) and received from the peer otherwise; results equal a plain load."""
    import torch
    import torch.distributed as dist

    from megatron_b200.core import dist_checkpointing as dc
    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor
    from megatron_b200.core.dist_checkpointing.strategies.fully_parallel import FullyParallelLoadStrategyWrapper

    tp, dp = 2, world // 2
    tp_rank, dp_rank = rank % tp, rank // tp
    dp_groups = [dist.new_group([t + tp * d for d in range(dp)]) for t in range(tp)]
    names = [f"layer{i}.w" for i in range(6)]
    fulls = {n: torch.arange(32 * 16, dtype=torch.float32).reshape(32, 16) * (i + 1) for i, n in enumerate(names)}

    def make(zero):
        return {n: ShardedTensor.from_rank_offsets(n, (torch.zeros(16, 16) if zero else fulls[n].chunk(tp, 0)[tp_rank].clone()), (0, tp_rank, tp), replica_id=dp_rank) for n in names}

    dc.save(make(False), ckpt_dir)
    for algo in ("broadcast", "gather_object"):
        strat = FullyParallelLoadStrategyWrapper(parallelization_group=dp_groups[tp_rank], exchange_algo=algo)
        out = dc.load(make(True), ckpt_dir, sharded_strategy=strat)
        for n in names:
            assert torch.equal(out[n], fulls[n].chunk(tp, 0)[tp_rank]), (algo, n)
        st = strat.last_stats
        assert st["shards_read"] == 3 and st["shards_received"] == 3, (algo, st)      # 6 shards per rank, half read, half received
    return True


def test_fully_parallel_load_reads_each_shard_once(tmp_path):
    from dist_utils import run_distributed

    d = tmp_path / "ck"
    d.mkdir()
    assert all(run_distributed(_fully_parallel_load_worker, 4, str(d)))


def _cached_plan_worker(rank, world, base):
    """Three saves with `cached_structure`: the first plans collectively (miss), the second reuses the plan (hit) and still writes the NEW values, a save with a
    changed structure on one rank falls back to full planning on every rank (miss) — and each checkpoint loads back correctly (sync and async-process paths)."""
    import os

    import torch

    from megatron_b200.core import dist_checkpointing as dc
    from megatron_b200.core.dist_checkpointing.mapping import ShardedObject, ShardedTensor
    from megatron_b200.core.dist_checkpointing.strategies import torch_dist
    from megatron_b200.core.dist_checkpointing.strategies.async_utils import AsyncCallsQueue, PersistentWriterProcess

    def sd(step, extra=False):
        w = torch.full((4, 6), float(step * 10 + rank))
        out = {"w": ShardedTensor.from_rank_offsets("w", w, (0, rank, world)), "r": ShardedTensor.from_rank_offsets("r", torch.full((3,), float(step)), replica_id=rank),
               "obj": ShardedObject("obj", {"step": step}, (1,), (0,), replica_id=rank), "iteration": step}
        if extra:
            out["new"] = ShardedTensor.from_rank_offsets("new", torch.full((2,), float(rank)), (0, rank, world))
        return out

    cache = torch_dist.get_plan_cache(None)
    dirs = []
    for step, extra, use_async in [(1, False, False), (2, False, False), (3, False, True), (4, True, False)]:
        d = os.path.join(base, f"s{step}")
        if rank == 0:
            os.makedirs(d)
        torch.distributed.barrier()
        req = dc.save(sd(step, extra), d, cached_structure=True, async_sharded_save=use_async)
        if use_async:
            q = AsyncCallsQueue()
            q.schedule_async_request(req)
            q.maybe_finalize_async_calls(blocking=True)
        dirs.append((step, extra, d))
    assert (cache.misses, cache.hits) == (2, 2), (cache.misses, cache.hits)
    for step, extra, d in dirs:
        tmpl = {"w": ShardedTensor.from_rank_offsets("w", torch.zeros(4, 6), (0, rank, world)), "r": ShardedTensor.from_rank_offsets("r", torch.zeros(3), replica_id=rank),
                "obj": ShardedObject("obj", None, (1,), (0,), replica_id=rank)}
        if extra:
            tmpl["new"] = ShardedTensor.from_rank_offsets("new", torch.zeros(2), (0, rank, world))
        got = dc.load(tmpl, d)
        assert torch.equal(got["w"], torch.full((4, 6), float(step * 10 + rank))) and torch.equal(got["r"], torch.full((3,), float(step))) and got["obj"] == {"step": step}
        assert got["iteration"] == step and (not extra or torch.equal(got["new"], torch.full((2,), float(rank))))
    PersistentWriterProcess.get().close()
    return True


def test_cached_save_plans(tmp_path):
    from dist_utils import run_distributed

    assert all(run_distributed(_cached_plan_worker, 2, str(tmp_path)))


def _tensor_aware_worker(rank, world):
    """Local-checkpoint container: pop → (pickle the hollow skeleton, ship the tensors) → refill → state dict; `fully_parallel` keeps DP-replicated tensors on
    the first rank of the group only and re-broadcasts them, rank-private shards stay where they are."""
    import pickle

    import torch

    from megatron_b200.core.dist_checkpointing.mapping import ShardedObject, ShardedTensor
    from megatron_b200.core.dist_checkpointing.tensor_aware_state_dict import MCoreTensorAwareStateDict as TASD

    def gen(fill):
        return {"model": {"repl": ShardedTensor.from_rank_offsets("repl", torch.full((4, 3), fill), replica_id=(0, 0, rank)),
                          "mine": ShardedTensor.from_rank_offsets("mine", torch.full((2,), fill + rank), (0, rank, world), replica_id=(0, 0, 0))},
                "obj": ShardedObject("o", {"k": fill}, (1,), (0,), replica_id=rank), "iteration": 7}

    for algo in ("atomic", "fully_parallel"):
        t = TASD.from_state_dict(gen(5.0), algo=algo, parallelization_group=None)
        n_stored = len(list(t.tensors))
        assert n_stored == (2 if algo == "atomic" or rank == 0 else 1), (algo, rank, n_stored)
        tensors = t.pop_tensors()
        assert t.is_hollow
        t2 = pickle.loads(pickle.dumps(t))
        t2.init_tensors()
        t2.insert_tensors([x.clone() for x in tensors])
        t2.copy_tensors_to_cpu()
        t2.restore_tensor_device()
        sd = t2.to_state_dict(gen(0.0))
        assert torch.equal(sd["model"]["repl"], torch.full((4, 3), 5.0)) and torch.equal(sd["model"]["mine"], torch.full((2,), 5.0 + rank))
        assert sd["obj"] == {"k": 5.0} and sd["iteration"] == 7
    try:
        TASD.from_state_dict(gen(1.0), algo="two_stage")
    except NotImplementedError:
        return True
    return False


def test_tensor_aware_state_dict_local_checkpoint_container():
    from dist_utils import run_distributed

    assert all(run_distributed(_tensor_aware_worker, 2))


def _ring_fn_worker(rank, world):
    """_RingAttnFn over 4 ranks (zig-zag chunks, GQA, manual ring backward with travelling dK/dV) vs plain attention on the gathered sequence."""
    import torch.distributed as dist

    from megatron_b200.ops import reference as ref
    from megatron_b200.parallel.context_parallel import _RingAttnFn, _RingShift

    torch.manual_seed(5)
    cp, c, b, hq, hk, d = world, 6, 2, 4, 2, 8
    s = 2 * cp * c
    full = [torch.randn(s, b, h, d) for h in (hq, hk, hk)]
    go_full = torch.randn(s, b, hq, d)
    idx = torch.cat([torch.arange(rank * c, (rank + 1) * c), torch.arange((2 * cp - 1 - rank) * c, (2 * cp - rank) * c)])
    q, k, v = (t[idx].clone().requires_grad_(True) for t in full)
    group = dist.group.WORLD
    for causal in (True, False):
        for t in (q, k, v):
            t.grad = None
        from megatron_b200.parallel.context_parallel import _AsyncRing

        out = _RingAttnFn.apply(q, k, v, 0.3, causal, rank, cp, _AsyncRing(group) if causal else (lambda x, reverse: _RingShift._shift(x, group, reverse)))
        out.backward(go_full[idx])
        fq, fk, fv = (t.clone().requires_grad_(True) for t in full)
        want = ref.attention_fwd(fq, fk, fv, causal, 0.3)
        want.backward(go_full)
        assert torch.allclose(out, want[idx], atol=1e-5), (causal, (out - want[idx]).abs().max())
        for got, w in ((q.grad, fq.grad), (k.grad, fk.grad), (v.grad, fv.grad)):
            assert torch.allclose(got, w[idx], atol=1e-5), (causal, (got - w[idx]).abs().max())
    return True


def test_ring_attention_function_four_ranks():
    from dist_utils import run_distributed

    assert all(run_distributed(_ring_fn_worker, 4))
