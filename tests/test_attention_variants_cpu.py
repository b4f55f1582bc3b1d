"""Absorbed MLA (latent-space attention) and DeepSeek sparse attention."""
import pytest
import torch

from dist_utils import run_distributed

_KW = dict(use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)


def _mla_cfg(**kw):
    from megatron_b200.core.transformer.transformer_config import MLATransformerConfig

    base = dict(num_layers=2, hidden_size=64, num_attention_heads=4, q_lora_rank=16, kv_lora_rank=24, qk_head_dim=16, qk_pos_emb_head_dim=8, v_head_dim=16,
                ffn_hidden_size=128, gated_linear_unit=True, activation_func=torch.nn.functional.silu, add_bias_linear=False, rotary_scaling_factor=4.0,
                original_max_position_embeddings=16, mscale_all_dim=1.0, qk_layernorm=True, **_KW)
    base.update(kw)
    return MLATransformerConfig(**base)


def _absorbed(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference_params import InferenceParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.experimental_attention_variant import AbsorbedMLASelfAttention

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    cfg = _mla_cfg()
    spec = get_gpt_layer_local_spec(multi_latent_attention=True, qk_layernorm=True, normalization="RMSNorm")
    torch.manual_seed(3)
    ref = GPTModel(cfg, spec, vocab_size=128, max_sequence_length=64, position_embedding_type="none")
    spec2 = get_gpt_layer_local_spec(multi_latent_attention=True, qk_layernorm=True, normalization="RMSNorm")
    spec2.submodules.self_attention.module = AbsorbedMLASelfAttention
    m = GPTModel(cfg, spec2, vocab_size=128, max_sequence_length=64, position_embedding_type="none")
    m.load_state_dict(ref.state_dict())                       # same parameters, same checkpoints
    assert isinstance(m.decoder.layers[0].self_attention, AbsorbedMLASelfAttention)
    tok = torch.randint(0, 128, (2, 32))
    pos = torch.arange(32)[None].expand(2, -1)
    l0 = ref(tok, pos, None, labels=tok).mean()
    l1 = m(tok, pos, None, labels=tok).mean()
    assert abs(l0.item() - l1.item()) < 1e-5
    l0.backward(), l1.backward()
    for (n, p), (_, q) in zip(ref.named_parameters(), m.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=2e-5, rtol=1e-3), n
    # decode on the latent cache: 32 = kv_lora_rank + rope dims per token and layer
    m.eval()
    with torch.no_grad():
        full = m(tok, pos, None)
        ip = InferenceParams(2, 64)
        outs = [m(tok[:, :20], pos[:, :20], None, inference_context=ip)]
        ip.sequence_len_offset = 20
        for t in range(20, 32):
            outs.append(m(tok[:, t : t + 1], pos[:, t : t + 1], None, inference_context=ip))
            ip.sequence_len_offset += 1
        assert (torch.cat(outs, 1) - full).abs().max().item() < 1e-4
        assert ip.key_value_memory_dict[1].shape[-1] == 24 + 8
    return True


def test_absorbed_mla_matches_mla_and_decodes_on_latents():
    assert run_distributed(_absorbed, 1) == [True]


def test_dsa_helpers_topk_sparse_attention_and_loss():
    from megatron_b200.core.transformer.experimental_attention_variant import dsa

    # layer schedule: offset 2, every 3rd layer computes
    comp = [dsa.source_dsa_compute_layer(l, 2, 3) for l in range(1, 10)]
    assert comp == [1, 2, 3, 3, 3, 6, 6, 6, 9] and not dsa.is_dsa_skip_topk_layer(6, 2, 3) and dsa.is_dsa_skip_topk_layer(7, 2, 3)
    x = torch.randn(5, 3, 16)
    assert torch.allclose(dsa.rotate_activation(dsa.rotate_activation(x)), x, atol=1e-5)          # normalised Hadamard is an involution
    torch.manual_seed(0)
    sq, b, h, d, n, c, r = 12, 2, 3, 8, 4, 10, 6
    q, w, k = torch.randn(sq, b, h, d), torch.rand(sq, b, h), torch.randn(sq, b, d)
    sc = dsa.compute_index_scores(q, w, k)
    ref = torch.stack([torch.stack([sum(w[t, bb, hh] * torch.relu(q[t, bb, hh] @ k[:, bb].T) for hh in range(h)) for t in range(sq)]) for bb in range(b)])
    assert torch.allclose(sc, ref, atol=1e-5)
    idx, valid = dsa.topk_causal_indices(sc, 4)
    for bb in range(b):
        for t in range(sq):
            sel = idx[bb, t][valid[bb, t]].tolist()
            assert all(s <= t for s in sel) and len(sel) == min(4, t + 1)
            assert torch.allclose(sc[bb, t, sel].sort().values, torch.topk(sc[bb, t, : t + 1], len(sel)).values.sort().values)     # ties (relu zeros) may pick either key
    # with k >= sequence length the sparse attention equals dense causal attention
    q_abs, kv = torch.randn(sq, b, n, c), torch.randn(sq, b, c)
    idx_all, valid_all = dsa.topk_causal_indices(sc, sq)
    out = dsa.sparse_attention_topk(q_abs, kv, idx_all, valid_all, 0.3, r)
    att = torch.einsum("sbnc,tbc->bnst", q_abs, kv) * 0.3
    att = att.masked_fill(torch.ones(sq, sq, dtype=torch.bool).triu(1)[None, None], float("-inf")).softmax(-1)
    dense = torch.einsum("bnst,tbr->sbnr", att, kv[..., :r])
    assert torch.allclose(out, dense, atol=1e-5)
    # the KL loss is zero when the index scores reproduce log of the (head-summed) attention distribution, positive otherwise
    tgt = att.sum(1)
    tgt = tgt / tgt.sum(-1, keepdim=True)
    perfect = torch.log(tgt.clamp(min=1e-30))
    assert dsa.compute_dsa_indexer_loss(perfect, q_abs, kv, 0.3, 1.0).item() < 1e-5
    assert dsa.compute_dsa_indexer_loss(sc, q_abs, kv, 0.3, 1.0).item() > 1e-3
    assert dsa.compute_dsa_indexer_loss(sc, q_abs, kv, 0.3, 1.0, idx, valid).item() >= 0


def test_dsattention_trains_indexer_only_through_kl():
    from types import SimpleNamespace

    from megatron_b200.core.transformer.experimental_attention_variant import DSAttention

    torch.manual_seed(1)
    cfg = SimpleNamespace(hidden_size=32, use_cpu_initialization=True, params_dtype=torch.float32, layernorm_epsilon=1e-5, init_method=lambda w: torch.nn.init.normal_(w, std=0.1),
                          dsa_indexer_n_heads=2, dsa_indexer_head_dim=8, dsa_indexer_topk=4, dsa_indexer_loss_coeff=0.5, dsa_topk_freq=2, dsa_skip_topk_offset=0)
    a1 = DSAttention(cfg, layer_number=1, softmax_scale=0.25)
    a2 = DSAttention(cfg, layer_number=2, softmax_scale=0.25)
    assert a1.indexer is not None and a2.indexer is None
    hs = torch.randn(10, 2, 32, requires_grad=True)
    q_abs = torch.randn(10, 2, 3, 12, requires_grad=True)
    kv = torch.randn(10, 2, 12, requires_grad=True)
    a1.train()
    out = a1(q_abs, kv, hs, v_width=8)
    assert out.shape == (10, 2, 3, 8)
    out2 = a2(q_abs, kv, hs, v_width=8, shared_indices=a1.last_indices)
    (out.sum() + out2.sum()).backward()
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in a1.indexer.parameters())      # reached through the KL auto-scaler
    assert hs.grad is None or hs.grad.abs().sum() == 0                                               # the index branch is detached from the trunk
    assert q_abs.grad.abs().sum() > 0 and kv.grad.abs().sum() > 0


def _spec(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.data_parallel_coordinator import DataParallelInferenceCoordinator
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.inference.speculative import SpeculativeDecoder, verify_draft_tokens
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()

    def build(seed, layers):
        torch.manual_seed(seed)
        cfg = TransformerConfig(num_layers=layers, hidden_size=64, num_attention_heads=4, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                                add_bias_linear=False, normalization="RMSNorm", **_KW)
        return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope")

    target, draft = build(1, 2), build(2, 1)
    prompt = [5, 17, 3, 42, 8]
    greedy = StaticInferenceEngine(target, max_sequence_length=128).generate([prompt], SamplingParams(temperature=0.0, num_tokens_to_generate=24))[0]
    # a different (bad) draft: few acceptances, identical output
    sd = SpeculativeDecoder(target, draft, num_speculative_tokens=3, max_sequence_length=128)
    assert sd.generate(prompt, SamplingParams(temperature=0.0, num_tokens_to_generate=24)) == greedy
    assert sd.stats.proposed > 0 and sd.stats.accepted <= sd.stats.proposed
    # the target as its own draft: everything is accepted, k + 1 tokens per target forward
    sd2 = SpeculativeDecoder(target, target, num_speculative_tokens=3, max_sequence_length=128)
    assert sd2.generate(prompt, SamplingParams(temperature=0.0, num_tokens_to_generate=24)) == greedy
    assert sd2.stats.acceptance_rate == 1.0 and sd2.stats.tokens_per_target_forward > 2.5
    # rejection rule is distribution preserving: empirical next-token distribution ≈ target distribution
    g = torch.Generator().manual_seed(0)
    pt = torch.tensor([[0.1, 0.6, 0.3], [0.3, 0.3, 0.4]])
    pdr = torch.tensor([[0.5, 0.25, 0.25]])
    counts = torch.zeros(3)
    for _ in range(4000):
        d = torch.multinomial(pdr[0], 1, generator=g)
        n, nxt = verify_draft_tokens(d, pdr, pt, g)
        counts[int(d) if n == 1 else nxt] += 1
    assert (counts / 4000 - pt[0]).abs().max() < 0.03
    # DP coordinator: least-loaded routing over two continuous-batching engines, results identical to a single engine
    e = [DynamicInferenceEngine(target, num_blocks=64, block_size=8, max_running=4, vocab_size=96) for _ in range(2)]
    coord = DataParallelInferenceCoordinator(e)
    prompts = [[5, 17, 3], [9, 9, 1, 4, 7, 7], [2], [30, 31, 32, 33]]
    sp = SamplingParams(temperature=0.0, num_tokens_to_generate=6)
    gids = [coord.add_request(p, sp) for p in prompts]
    assert sorted(coord.placement.values()) == [0, 0, 1, 1]
    res = coord.run_until_done()
    ref = StaticInferenceEngine(target, max_sequence_length=128)
    for gid, p in zip(gids, prompts):
        assert res[gid].generated_tokens == ref.generate([p], sp)[0]
    assert all(r.outstanding_tokens == 0 for r in coord.replicas)
    coord.pause(0)
    assert coord.placement[coord.add_request([1, 2], sp)] == 1
    return True


def test_speculative_decoding_and_dp_coordinator():
    assert run_distributed(_spec, 1) == [True]


def test_symmetric_memory_manager_fallback():
    from megatron_b200.core.inference.symmetric_memory import SymmetricMemoryManager

    m = SymmetricMemoryManager(max_bytes=1 << 12)
    a = m.get_buffer("ar", (4, 8), torch.float32, device="cpu")
    assert m.get_buffer("ar", (4, 8), torch.float32) is a and not a.is_symmetric
    a.tensor.fill_(2.0)
    assert a.all_reduce_().sum() == 64.0
    import pytest

    with pytest.raises(MemoryError):
        m.get_buffer("big", (1 << 12,), torch.float32, device="cpu")


def _disagg(rank, world):
    import torch.distributed as dist
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.disaggregation import DecodeWorker, InProcessTransport, PrefillWorker, TorchDistTransport
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()            # tp = 1: every rank holds a full replica; rank 0 = prefill pool, rank 1 = decode pool
    torch.manual_seed(1)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                            add_bias_linear=False, normalization="RMSNorm", **_KW)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope")
    prompts = [[5, 17, 3, 42, 8, 1, 2], [9, 9], [30, 31, 32, 33, 34]]
    sp = SamplingParams(temperature=0.0, num_tokens_to_generate=8, stop_token_ids=(95,))
    ref = StaticInferenceEngine(model, max_sequence_length=128)
    expected = [ref.generate([p], sp)[0] for p in prompts]
    mk = lambda: DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=4, vocab_size=96)  # noqa: E731
    # same process
    pre, dec, tr = PrefillWorker(mk()), DecodeWorker(mk()), InProcessTransport()
    for i, p in enumerate(prompts):
        tr.send(pre.prefill(i, p, sp))
    assert pre.engine.cache.allocator.num_free == 64           # prefill pages are handed over, not kept
    while (pl := tr.recv()) is not None:
        assert dec.admit(pl)
    out = dec.run_until_done()
    assert [out[i].generated_tokens for i in range(3)] == expected and tr.bytes_moved > 0
    # across ranks over torch.distributed
    t = TorchDistTransport()
    if rank == 0:
        w = PrefillWorker(mk())
        for i, p in enumerate(prompts):
            t.send(w.prefill(i, p, sp), dst=1)
        res = None
    else:
        w = DecodeWorker(mk())
        for _ in prompts:
            assert w.admit(t.recv(src=0))
        fin = w.run_until_done()
        res = [fin[i].generated_tokens for i in range(3)]
        assert res == expected
    dist.barrier()
    return res


def test_prefill_decode_disaggregation_matches_single_engine():
    res = run_distributed(_disagg, 2)
    assert res[0] is None and res[1] is not None


def _batched_decode(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    torch.manual_seed(3)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=8, num_query_groups=2, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                            add_bias_linear=False, normalization="RMSNorm", **_KW)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope")
    prompts = [[5, 17, 3, 42, 8, 1, 2, 7, 7, 7, 11], [9, 9], [30, 31, 32, 33, 34], [1], [2, 3, 4, 5, 6, 7, 8, 9, 10]]
    gens = [6, 11, 3, 9, 7]                                     # requests leave at different steps; the 5th is admitted late (max_running 4)
    ref = StaticInferenceEngine(model, max_sequence_length=128)
    want = [ref.generate([p], SamplingParams(temperature=0.0, num_tokens_to_generate=n))[0] for p, n in zip(prompts, gens)]
    out = {}
    for batched in (True, False):
        e = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=4, vocab_size=96, batched_decode=batched)
        ids = [e.add_request(p, SamplingParams(temperature=0.0, num_tokens_to_generate=n, return_log_probs=True)) for p, n in zip(prompts, gens)]
        fin = e.run_until_done()
        out[batched] = ([fin[i].generated_tokens for i in ids], [fin[i].log_probs for i in ids], e.decode_forwards, e.steps)
        assert e.cache.allocator.num_free == 64
    assert out[True][0] == out[False][0] == want
    for a, b in zip(out[True][1], out[False][1]):
        assert torch.allclose(torch.tensor(a), torch.tensor(b), atol=1e-4)
    assert e.batched_decode is False and DynamicInferenceEngine(model, num_blocks=8, block_size=4).batched_decode is True       # auto-detected for standard attention
    # one forward per step instead of one per running request
    assert 0 < out[True][2] <= out[True][3] and out[False][2] == 0
    # bucketed (static-shape) decode: batch padded to {2, 4}, attended length to multiples of 16 → same tokens, few distinct shapes
    e = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=4, vocab_size=96, decode_batch_buckets=[2, 4])
    ids = [e.add_request(p, SamplingParams(temperature=0.0, num_tokens_to_generate=n)) for p, n in zip(prompts, gens)]
    fin = e.run_until_done()
    assert [fin[i].generated_tokens for i in ids] == want
    assert {b for b, _ in e.decode_shapes_seen} <= {2, 4} and all(L % 16 == 0 for _, L in e.decode_shapes_seen) and len(e.decode_shapes_seen) <= 4
    assert e.cache.allocator.num_free == 63                      # everything released except the scratch block
    # chunked prefill: at most 4 prompt tokens per step — the 11-token prompt takes 3 steps, during which the already running requests keep decoding;
    # tokens are identical, and the first token of the long prompt comes out only with its last chunk
    e = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=4, vocab_size=96, max_prefill_tokens_per_step=4)
    order = [1, 3, 0, 2, 4]                                       # short prompts first so that decodes are running while the long one prefills
    ids = {i: e.add_request(prompts[i], SamplingParams(temperature=0.0, num_tokens_to_generate=gens[i])) for i in order}
    seen_prefilling_while_decoding = False
    while e.has_unfinished():
        e.step()
        pre = [r for r in e.running if r.status == "prefilling"]
        if pre and e.decode_forwards > 0:
            seen_prefilling_while_decoding = True
            assert all(not r.generated_tokens and 0 < r.prefill_pos < len(r.prompt_tokens) for r in pre)
    assert [e.finished[ids[i]].generated_tokens for i in range(5)] == want
    assert seen_prefilling_while_decoding and e.prefill_chunks >= 3 + 1 + 2 + 1 + 3 - 2 and e.prefill_tokens == sum(len(p) for p in prompts)
    assert e.cache.allocator.num_free == 64
    return True


def test_batched_paged_decode_matches_per_request_decode():
    run_distributed(_batched_decode, 1)


def _prefix_cache(rank, world):
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.kv_cache import KVBlockAllocator
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    # allocator: pin / park / evict
    al = KVBlockAllocator(4, enable_prefix_caching=True)
    hs = al.chain_hashes(list(range(10)), 4)
    assert len(hs) == 2 and hs != al.chain_hashes([9] + list(range(1, 10)), 4) and hs[0] == al.chain_hashes(list(range(4)) + [7, 7, 7, 7], 4)[0]
    blocks = al.allocate(3)
    al.register(blocks[0], hs[0]), al.register(blocks[1], hs[1])
    al.release(blocks)
    assert al.num_free == 4 and len(al.parked) == 2                     # registered blocks are parked, not freed
    assert al.lookup_and_pin(hs) == blocks[:2] and al.num_free == 2
    al.release(blocks[:2])
    got = al.allocate(4)                                                # needs the parked ones → evicts them, oldest first
    assert sorted(got) == [0, 1, 2, 3] and al.evictions == 2 and al.lookup_and_pin(hs) == []

    ps.initialize_model_parallel()
    torch.manual_seed(4)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                            add_bias_linear=False, normalization="RMSNorm", **_KW)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope")
    system = [7, 3, 9, 1, 4, 4, 8, 2, 6, 6, 5, 1, 3]                     # 13 tokens = 3 full blocks of 4 + 1
    prompts = [system + [20, 21], system + [30], system[:8] + [40, 41, 42], system]
    sp = SamplingParams(temperature=0.0, num_tokens_to_generate=5)
    plain = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=1, vocab_size=96)
    ids = [plain.add_request(p, sp) for p in prompts]
    want = [plain.run_until_done()[i].generated_tokens for i in ids]
    eng = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=1, vocab_size=96, enable_prefix_caching=True)
    ids = [eng.add_request(p, sp) for p in prompts]
    fin = eng.run_until_done()
    assert [fin[i].generated_tokens for i in ids] == want
    # request 0 computes everything; 1 reuses 12 tokens, 2 reuses 8, 3 (identical to the system prompt, 13 tokens) reuses 12
    assert plain.prefill_tokens == sum(map(len, prompts)) and eng.prefill_tokens == 15 + (14 - 12) + (11 - 8) + (13 - 12)
    assert eng.cache.allocator.num_free == 64 and eng.cache.allocator.hits == 3 + 2 + 3
    # concurrent sharing: both requests hold the same physical blocks while running
    eng2 = DynamicInferenceEngine(model, num_blocks=64, block_size=4, max_running=4, vocab_size=96, enable_prefix_caching=True)
    a = eng2.add_request(prompts[0], sp)
    eng2.step()
    b = eng2.add_request(prompts[1], sp)
    eng2.step()
    assert eng2.cache.block_tables[a][:3] == eng2.cache.block_tables[b][:3] and eng2.cache.allocator.ref[eng2.cache.block_tables[a][0]] == 2
    fin2 = eng2.run_until_done()
    assert [fin2[a].generated_tokens, fin2[b].generated_tokens] == want[:2]
    return True


def test_prefix_caching_reuses_blocks_and_matches_uncached_generation():
    run_distributed(_prefix_cache, 1)


def _zmq_serving(rank, world):
    import socket
    import time

    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.inference.zmq_coordinator import ZMQInferenceClient, start_in_threads
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()

    def build():
        torch.manual_seed(3)
        cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=8, num_query_groups=2, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                                add_bias_linear=False, normalization="RMSNorm", **_KW)
        return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope").eval()

    models = [build(), build()]                                   # two data-parallel replicas (same weights), one engine + one worker thread each
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    engines = [DynamicInferenceEngine(m, num_blocks=64, block_size=4, max_running=4, vocab_size=96) for m in models]
    coord_thread, workers = start_in_threads(engines, port)
    client = ZMQInferenceClient(port)
    prompts = [[5, 17, 3, 42, 8, 1, 2, 7], [9, 9], [30, 31, 32, 33, 34], [1], [2, 3, 4, 5, 6, 7, 8, 9, 10], [4, 4, 4]]
    sp = SamplingParams(temperature=0.0, num_tokens_to_generate=6, return_log_probs=True)
    ref = StaticInferenceEngine(build(), max_sequence_length=128)
    want = [ref.generate([p], SamplingParams(temperature=0.0, num_tokens_to_generate=6))[0] for p in prompts]
    got = client.generate(prompts, sp)
    assert got == want
    st = client.stats()
    assert sum(st["served"]) == 6 and min(st["served"]) >= 1 and st["outstanding_tokens"] == [0, 0]      # both replicas took work, nothing left outstanding
    assert all(len(client.results[i]["log_probs"]) == 6 and client.results[i]["ttft"] is not None for i in range(6))
    # pause: engines acknowledge and stop stepping; submitted work waits until unpause
    client.pause_engines()
    t0 = time.time()
    while client.stats()["paused_acks"] < 2 and time.time() - t0 < 10:
        time.sleep(0.02)
    rid = client.submit([7, 8, 9], sp)
    time.sleep(0.3)
    assert rid not in client.results and not client.sock.poll(0)
    client.unpause_engines()
    assert client.collect([rid])[rid]["generated_tokens"] == ref.generate([[7, 8, 9]], SamplingParams(temperature=0.0, num_tokens_to_generate=6))[0]
    client.stop()
    coord_thread.join(10)
    for w in workers:
        w.join(10)
    assert not coord_thread.is_alive() and not any(w.is_alive() for w in workers)
    return True


def test_zmq_coordinator_routes_between_data_parallel_engines():
    run_distributed(_zmq_serving, 1)


def _tp2_serving(rank, world):
    """The dynamic engine on a TP=2 model (2 gloo ranks, each with its half of the KV heads in its own paged cache): every rank runs the same schedule and
    emits the same tokens as the TP=1 static engine — incl. chunked prefill and bucketed decode."""
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    def build(tp):
        cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=8, num_query_groups=4, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                                add_bias_linear=False, normalization="RMSNorm", tensor_model_parallel_size=tp, **_KW)
        # serving models gather the vocabulary-parallel logits (parallel_output=False): every TP rank samples from the full distribution
        return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope", parallel_output=False).eval()

    # TP=1 reference weights, built identically on every rank, then sharded by hand into the TP=2 model
    ps.initialize_model_parallel(tensor_model_parallel_size=1)
    model_parallel_cuda_manual_seed(3)
    torch.manual_seed(3)
    full = build(1)
    prompts = [[5, 17, 3, 42, 8, 1, 2, 7, 7, 7, 11], [9, 9], [30, 31, 32, 33, 34], [1]]
    gens = [6, 9, 3, 7]
    ref = StaticInferenceEngine(full, max_sequence_length=128)
    want = [ref.generate([p], SamplingParams(temperature=0.0, num_tokens_to_generate=n))[0] for p, n in zip(prompts, gens)]
    sd = {k: v.clone() for k, v in full.state_dict().items() if isinstance(v, torch.Tensor)}
    ps.destroy_model_parallel()
    ps.initialize_model_parallel(tensor_model_parallel_size=2)
    model_parallel_cuda_manual_seed(3)
    tp_model = build(2)
    with torch.no_grad():
        for name, p in tp_model.named_parameters():
            w = sd[name]
            if p.shape == w.shape:
                p.copy_(w)
            elif name.endswith("linear_qkv.weight"):                    # [groups x (q per group + k + v) x d, h]: split by query group
                p.copy_(w.view(4, -1, w.shape[-1]).chunk(2, 0)[rank].reshape(p.shape))
            elif name.endswith("linear_fc1.weight"):                    # [gate; up] each split over TP
                g, u = w.chunk(2, 0)
                p.copy_(torch.cat([g.chunk(2, 0)[rank], u.chunk(2, 0)[rank]], 0))
            else:
                dim = 0 if p.shape[0] != w.shape[0] else 1
                p.copy_(w.chunk(2, dim)[rank])
    tp_model.output_layer.gather_output, tp_model.parallel_output = False, True      # a training-style model: the engine switches it to gathered logits itself
    for kw in (dict(), dict(max_prefill_tokens_per_step=4), dict(decode_batch_buckets=[2, 4])):
        e = DynamicInferenceEngine(tp_model, num_blocks=64, block_size=4, max_running=4, vocab_size=96, **kw)
        assert tp_model.output_layer.gather_output is True
        assert e.cache.k.shape[3] == 2                                  # this rank's half of the 4 KV heads
        ids = [e.add_request(p, SamplingParams(temperature=0.0, num_tokens_to_generate=n)) for p, n in zip(prompts, gens)]
        fin = e.run_until_done()
        assert [fin[i].generated_tokens for i in ids] == want, (rank, kw)
    return True


def test_dynamic_engine_on_tensor_parallel_model():
    assert all(run_distributed(_tp2_serving, 2))


def _variant_specs(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.experimental_attention_variant_module_specs import (get_linear_attention_pattern, get_moe_layer_pattern,
                                                                                          get_transformer_block_with_experimental_attention_variant_spec,
                                                                                          normalize_experimental_attention_variant)
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import MLATransformerConfig, TransformerConfig

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    tok = torch.randint(0, 64, (2, 32), generator=torch.Generator().manual_seed(0))
    pos = torch.arange(32).unsqueeze(0).expand(2, -1)
    # linear attention: gated-delta-net mixers in the attention slot of 3 layers out of 4 (freq 4), softmax attention in the fourth; MoE in every other layer
    cfg = TransformerConfig(num_layers=4, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0,
                            experimental_attention_variant="gdn", linear_attention_freq=4, linear_key_head_dim=16, linear_value_head_dim=16, linear_num_key_heads=2,
                            linear_num_value_heads=4, num_moe_experts=4, moe_router_topk=2, moe_layer_freq=2, moe_ffn_hidden_size=32)
    assert get_linear_attention_pattern(cfg) == [1, 1, 1, 0] and get_moe_layer_pattern(cfg) == [1, 0, 1, 0]
    with pytest.warns(DeprecationWarning):
        assert normalize_experimental_attention_variant("gated_delta_net") == "gdn"
    m = GPTModel(cfg, get_transformer_block_with_experimental_attention_variant_spec(cfg), vocab_size=64, max_sequence_length=32, position_embedding_type="rope")
    assert [type(l.self_attention).__name__ for l in m.decoder.layers] == ["GatedDeltaNetAttention"] * 3 + ["SelfAttention"]
    assert [type(l.mlp).__name__ for l in m.decoder.layers] == ["MoELayer", "MLP", "MoELayer", "MLP"]
    m(tok, pos, None, labels=tok).mean().backward()            # MoE experts with bias terms: the expert output bias is folded into the expert output
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # sparse attention: every layer is absorbed MLA with the DSA core; indexers in layers 1 and 3, layers 2 and 4 reuse their top-k
    mc = MLATransformerConfig(num_layers=4, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, q_lora_rank=32,
                              kv_lora_rank=32, qk_head_dim=16, qk_pos_emb_head_dim=8, v_head_dim=16, rope_type="rope", experimental_attention_variant="dsa",
                              dsa_indexer_n_heads=2, dsa_indexer_head_dim=16, dsa_indexer_topk=8, dsa_indexer_loss_coeff=0.1, dsa_indexer_topk_freq=2, add_bias_linear=False)
    m2 = GPTModel(mc, get_transformer_block_with_experimental_attention_variant_spec(mc), vocab_size=64, max_sequence_length=32, position_embedding_type="rope")
    assert [l.self_attention.dsa.indexer is not None for l in m2.decoder.layers] == [True, False, True, False]
    loss = m2(tok, pos, None, labels=tok).mean()
    loss.backward()
    assert all(float(p.grad.abs().sum()) > 0 for p in m2.decoder.layers[0].self_attention.dsa.parameters())      # trained through the KL term only
    # with top-k >= sequence length the sparse core sees every causal key: same loss as the dense absorbed-MLA model with the same weights
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.transformer.experimental_attention_variant import AbsorbedMLASelfAttention

    full = MLATransformerConfig(**{**{f.name: getattr(mc, f.name) for f in __import__("dataclasses").fields(mc) if f.init}, "dsa_indexer_topk": 64, "dsa_indexer_loss_coeff": 0.0})
    m3 = GPTModel(full, get_transformer_block_with_experimental_attention_variant_spec(full), vocab_size=64, max_sequence_length=32, position_embedding_type="rope")
    dense_spec = get_gpt_layer_local_spec(multi_latent_attention=True)
    dense_spec.submodules.self_attention.module = AbsorbedMLASelfAttention
    m4 = GPTModel(full, dense_spec, vocab_size=64, max_sequence_length=32, position_embedding_type="rope")
    missing = m4.load_state_dict({k: v for k, v in m3.state_dict().items() if ".dsa." not in k}, strict=False)
    assert not missing.missing_keys
    with torch.no_grad():
        assert abs(float(m3(tok, pos, None, labels=tok).mean()) - float(m4(tok, pos, None, labels=tok).mean())) < 1e-4
    return True


def test_experimental_attention_variant_block_specs():
    from dist_utils import run_distributed

    assert run_distributed(_variant_specs, 1) == [True]
