"""Command-line tools: parallel preprocessing, checkpoint inspector, routing analysis, per-dataset sequence counts, NCCL group options."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "checkpoint"))
sys.path.insert(0, os.path.join(ROOT, "tools", "moe_routing"))


def _corpus(path, n=200):
    rng = np.random.default_rng(0)
    docs = []
    with open(path, "w") as f:
        for i in range(n):
            ids = rng.integers(1, 90, size=int(rng.integers(1, 40))).tolist()
            docs.append(ids)
            f.write(json.dumps({"text": " ".join(map(str, ids)), "meta": i}) + "\n")
            if i % 17 == 0:
                f.write("\n")                                     # blank lines are skipped
    return docs


def test_preprocess_data_fast_keeps_order_and_matches_serial(tmp_path):
    import preprocess_data_fast as fast

    from megatron_b200.core.datasets.indexed_dataset import IndexedDataset

    src = tmp_path / "c.jsonl"
    docs = _corpus(src)
    ranges = fast.line_aligned_ranges(str(src), 7)
    assert ranges[0][0] == 0 and ranges[-1][1] == os.path.getsize(src) and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    with open(src, "rb") as f:
        for s, _ in ranges[1:]:
            f.seek(s - 1)
            assert f.read(1) == b"\n"                              # every range starts at a line start
    common = ["--input", str(src), "--tokenizer-type", "NullTokenizer", "--vocab-size", "100", "--append-eod"]
    n1 = fast.main(common + ["--output-prefix", str(tmp_path / "serial"), "--workers", "1", "--chunks-per-worker", "1"])
    n2 = fast.main(common + ["--output-prefix", str(tmp_path / "par"), "--workers", "3", "--chunks-per-worker", "2"])
    assert n1 == n2 == len(docs)
    a, b = IndexedDataset(str(tmp_path / "serial_text_document")), IndexedDataset(str(tmp_path / "par_text_document"))
    assert len(a) == len(b) == len(docs)
    for i in (0, 1, 57, 199):
        assert np.array_equal(a[i], b[i]) and a[i][:-1].tolist() == docs[i]
    assert np.array_equal(a.index.document_indices, b.index.document_indices)
    assert not [f for f in os.listdir(tmp_path) if ".part" in f]     # partial files are cleaned up

    import build_sequences_per_dataset as bs

    out = bs.main(["--data-path", "0.5", str(tmp_path / "serial_text_document"), "0.5", str(tmp_path / "par_text_document"),
                   "--per-dataset-sequences-path", str(tmp_path / "seq.json")])
    assert list(out.values()) == [(200, 200), (200, 200)]


def test_checkpoint_inspector_inspect_diff_rename(tmp_path, capsys):
    import checkpoint_inspector as ci
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint import FileSystemWriter

    a = {"decoder.layers.0.w": torch.arange(12.0).view(3, 4), "decoder.final.b": torch.ones(4, dtype=torch.bfloat16)}
    b = {"decoder.layers.0.w": a["decoder.layers.0.w"] + 0.5, "extra": torch.zeros(2)}
    dcp.save(a, storage_writer=FileSystemWriter(str(tmp_path / "a")), no_dist=True)
    dcp.save(b, storage_writer=FileSystemWriter(str(tmp_path / "b")), no_dist=True)
    meta = ci.main(["inspect", str(tmp_path / "a")])
    assert meta["decoder.layers.0.w"] == ((3, 4), torch.float32) and meta["decoder.final.b"][1] == torch.bfloat16
    assert "2 entries" in capsys.readouterr().out
    only_a, only_b, changed, worst = ci.main(["diff", str(tmp_path / "a"), str(tmp_path / "b"), "--values"])
    assert only_a == ["decoder.final.b"] and only_b == ["extra"] and not changed and abs(worst["decoder.layers.0.w"] - 0.5) < 1e-6
    ci.main(["rename", str(tmp_path / "a"), str(tmp_path / "c"), "--sub", r"^decoder\.", "encoder."])
    got = ci.load_full(str(tmp_path / "c"))
    assert set(got) == {"encoder.layers.0.w", "encoder.final.b"} and torch.equal(got["encoder.layers.0.w"], a["decoder.layers.0.w"])


def test_routing_analysis_from_tracer_records(tmp_path):
    import analyze_routing as ar

    from megatron_b200.core.transformer.moe.router_trace import RouterTracer

    tr = RouterTracer(str(tmp_path), flush_every=100)
    same = torch.tensor([[0, 1], [0, 1], [2, 3], [0, 1]])
    for step in range(3):
        tr.record_indices(same, module_name="decoder.layers.0.mlp.router")                     # skewed, perfectly repeatable
        tr.record_indices(torch.tensor([[step % 4, (step + 1) % 4]] * 4), module_name="decoder.layers.1.mlp.router")
        tr.advance_step()
    tr.flush()
    rep = ar.main([tr.trace_dir, "--num-experts", "4", "--json", str(tmp_path / "r.json")])
    l0, l1 = rep["decoder/0"], rep["decoder/1"]
    assert l0["repeat_fraction"] == 1.0 and l0["slots"] == 24 and abs(l0["load"][0] - 3 / 8) < 1e-6 and abs(l0["imbalance_max_over_mean"] - 1.5) < 1e-6
    assert l1["repeat_fraction"] == 0.0 and l0["entropy"] < 1.0 and l0["dead_experts"] == 0
    assert json.load(open(tmp_path / "r.json"))["decoder/1"]["records"] == 3


def test_nccl_communicator_config_parsing(tmp_path):
    from megatron_b200.core import parallel_state as ps

    p = tmp_path / "nccl.yaml"
    p.write_text("tp:\n  max_ctas: 8\n  cga_cluster_size: 2\n  is_high_priority_stream: true\ndp:\n  min_ctas: 2\ndefault:\n  max_ctas: 24\n")
    cfg = ps.load_nccl_communicator_config(str(p))
    tp = ps.get_nccl_options("tp", cfg)
    assert tp.is_high_priority_stream and tp.config.max_ctas == 8 and tp.config.cga_cluster_size == 2
    dp = ps.get_nccl_options("dp", cfg)
    assert dp.config.min_ctas == 2 and dp.config.max_ctas == 24 and not dp.is_high_priority_stream      # group fields over the defaults
    assert ps.get_nccl_options("pp", {}) is None and ps.get_nccl_options("pp", {}, high_priority=True).is_high_priority_stream
    p.write_text("tpp:\n  max_ctas: 8\n")
    with pytest.raises(ValueError):
        ps.load_nccl_communicator_config(str(p))
    p.write_text("tp:\n  max_cta: 8\n")
    with pytest.raises(ValueError):
        ps.load_nccl_communicator_config(str(p))
    p.write_text("tp:\n  net_name: tcp\n")
    with pytest.raises(RuntimeError):
        ps.load_nccl_communicator_config(str(p))


def test_trtllm_export_layout_split_and_distributed_variant(tmp_path):
    from megatron_b200.core.export.trtllm import (DistributedTRTLLMWeightsConverter, ExportConfig, TRTLLMLayers, TRTLLMWeightsConverter, pad_vocab_size,
                                                   save_trtllm_checkpoint, trtllm_model_config)
    from megatron_b200.core.export.hf_llama import megatron_to_hf_llama
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=8, num_query_groups=2, ffn_hidden_size=48, gated_linear_unit=True,
                            activation_func=torch.nn.functional.silu, add_bias_linear=False, use_cpu_initialization=True)
    h, d, g, r = 32, cfg.kv_channels, 2, 4
    sd = {"embedding.word_embeddings.weight": torch.randn(100, h), "output_layer.weight": torch.randn(100, h), "decoder.final_layernorm.weight": torch.randn(h)}
    for i in range(2):
        p = f"decoder.layers.{i}."
        sd.update({p + "self_attention.linear_qkv.weight": torch.randn(g * (r + 2) * d, h), p + "self_attention.linear_proj.weight": torch.randn(h, r * g * d),
                   p + "mlp.linear_fc1.weight": torch.randn(96, h), p + "mlp.linear_fc2.weight": torch.randn(h, 48), p + "input_layernorm.weight": torch.randn(h),
                   p + "pre_mlp_layernorm.weight": torch.randn(h), p + "self_attention.linear_qkv._extra_state": None})
    hf = megatron_to_hf_llama(sd, 8, 2, d)
    assert TRTLLMLayers.return_layer_name_and_number("decoder.layers.12.mlp.linear_fc1.weight") == ("decoder.layers.mlp.linear_fc1.weight", 12)
    for tp in (1, 2, 4):                                             # tp 4 > 2 KV groups → K/V heads replicated
        ex = ExportConfig(inference_tp_size=tp, dtype=torch.float32)
        ranks = TRTLLMWeightsConverter(ex, cfg).convert(sd, vocab_size=100)
        assert len(ranks) == tp
        qn, kn = 8 * d // tp, max(2 // tp, 1) * d
        qkv = [w["transformer.layers.1.attention.qkv.weight"] for w in ranks]
        assert all(t.shape == (qn + 2 * kn, h) for t in qkv)
        assert torch.equal(torch.cat([t[:qn] for t in qkv]), hf["model.layers.1.self_attn.q_proj.weight"])
        k_all = torch.cat([t[qn : qn + kn] for t in qkv])
        want_k = hf["model.layers.1.self_attn.k_proj.weight"].view(2, d, h).repeat_interleave(max(tp // 2, 1), 0).reshape(-1, h)
        assert torch.equal(k_all, want_k)
        assert torch.equal(torch.cat([w["transformer.layers.0.mlp.fc.weight"] for w in ranks]), hf["model.layers.0.mlp.gate_proj.weight"])
        assert torch.equal(torch.cat([w["transformer.layers.0.mlp.gate.weight"] for w in ranks]), hf["model.layers.0.mlp.up_proj.weight"])
        assert torch.equal(torch.cat([w["transformer.layers.0.mlp.proj.weight"] for w in ranks], 1), sd["decoder.layers.0.mlp.linear_fc2.weight"])
        assert torch.equal(torch.cat([w["transformer.layers.0.attention.dense.weight"] for w in ranks], 1), sd["decoder.layers.0.self_attention.linear_proj.weight"])
        emb = torch.cat([w["transformer.vocab_embedding.weight"] for w in ranks])
        assert emb.shape[0] == pad_vocab_size(100, tp) and torch.equal(emb[:100], sd["embedding.word_embeddings.weight"]) and float(emb[100:].abs().sum()) == 0
        assert all(torch.equal(w["transformer.ln_f.weight"], sd["decoder.final_layernorm.weight"]) for w in ranks)
    # distributed variant: rank r converts its own training shard and lands on the same tensors
    tp = 2
    ranks = TRTLLMWeightsConverter(ExportConfig(tp, dtype=torch.float32), cfg).convert(sd, vocab_size=128)
    for rk in range(tp):
        shard = {}
        for k, v in sd.items():
            if v is None:
                continue
            if "linear_qkv" in k or "word_embeddings" in k or "output_layer" in k:
                vv = torch.cat([v, v.new_zeros(128 - v.shape[0], h)]) if v.shape[0] == 100 else v
                shard[k] = vv.chunk(tp, 0)[rk]
            elif "linear_fc1" in k:
                shard[k] = torch.cat([v.chunk(2, 0)[0].chunk(tp, 0)[rk], v.chunk(2, 0)[1].chunk(tp, 0)[rk]])
            elif "linear_fc2" in k or "linear_proj" in k:
                shard[k] = v.chunk(tp, 1)[rk]
            else:
                shard[k] = v
        got = DistributedTRTLLMWeightsConverter(ExportConfig(tp, dtype=torch.float32), cfg, rk, tp).convert(shard, vocab_size=128)
        assert got.keys() == ranks[rk].keys()
        for k in got:
            assert torch.equal(got[k], ranks[rk][k]), k
    # MoE experts (grouped layout) are stacked per layer with [up ; gate] halves
    mcfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=8, num_query_groups=2, ffn_hidden_size=48, num_moe_experts=4, moe_ffn_hidden_size=16,
                             gated_linear_unit=True, activation_func=torch.nn.functional.silu, add_bias_linear=False, use_cpu_initialization=True)
    msd = {"decoder.layers.0.mlp.experts.weight1": torch.randn(4, 32, h), "decoder.layers.0.mlp.experts.weight2": torch.randn(4, h, 16),
           "decoder.layers.0.mlp.router.weight": torch.randn(4, h), "embedding.word_embeddings.weight": torch.randn(64, h)}
    mr = TRTLLMWeightsConverter(ExportConfig(2, dtype=torch.float32), mcfg).convert(msd)
    fc = torch.cat([w["transformer.layers.0.mlp.fc.weight"] for w in mr], 1)          # [E, 2 * (2 * 8), h] rank-major halves
    assert fc.shape == (4, 32, h) and torch.equal(mr[0]["transformer.layers.0.mlp.fc.weight"][:, :8], msd["decoder.layers.0.mlp.experts.weight1"][:, 16:24])
    assert torch.equal(torch.cat([w["transformer.layers.0.mlp.proj.weight"] for w in mr], 2), msd["decoder.layers.0.mlp.experts.weight2"])
    assert torch.equal(mr[1]["lm_head.weight"], mr[1]["transformer.vocab_embedding.weight"])          # tied when no output layer
    conf = trtllm_model_config(cfg, 100, 4096, ExportConfig(2))
    assert conf["vocab_size"] == 128 and conf["mapping"]["tp_size"] == 2 and conf["num_key_value_heads"] == 2 and conf["hidden_act"] == "swiglu"
    save_trtllm_checkpoint(str(tmp_path / "trt"), ranks, conf)
    assert os.path.exists(tmp_path / "trt" / "config.json") and len([f for f in os.listdir(tmp_path / "trt") if f.startswith("rank")]) == 2


def test_prepare_cache_then_pretrain_reuses_it(tmp_path):
    """tools/prepare_cache.py builds the dataset index caches with the training job's own arguments; the training run then finds every index in the cache
    (no new files) and trains on the real data."""
    import subprocess

    import prepare_cache
    import preprocess_data_fast as fast

    src = tmp_path / "c.jsonl"
    _corpus(src, n=400)
    fast.main(["--input", str(src), "--tokenizer-type", "NullTokenizer", "--vocab-size", "100", "--append-eod", "--output-prefix", str(tmp_path / "corp"), "--workers", "1"])
    cache = tmp_path / "cache"
    common = ["--num-layers", "2", "--hidden-size", "32", "--num-attention-heads", "4", "--seq-length", "32", "--max-position-embeddings", "32", "--micro-batch-size", "2",
              "--global-batch-size", "4", "--train-iters", "3", "--lr", "1e-3", "--tokenizer-type", "NullTokenizer", "--vocab-size", "100", "--data-path",
              str(tmp_path / "corp_text_document"), "--split", "90,5,5", "--data-cache-path", str(cache), "--eval-iters", "1", "--eval-interval", "100", "--seed", "7"]
    rep = prepare_cache.main(common)
    assert rep["requested_samples"] == {"train": 12, "valid": 4, "test": 4} and rep["built"]["train"] >= 12 and rep["cache_files"] >= 6
    before = sorted(os.listdir(cache))
    with pytest.raises(SystemExit):
        prepare_cache.main(common + ["--mock-data"])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29661", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py")] + common + ["--log-interval", "1", "--distributed-backend", "gloo", "--lr-decay-iters", "10"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "iteration        3/       3" in r.stdout
    assert sorted(os.listdir(cache)) == before                      # the job found everything it needed


def test_run_inference_performance_test_cpu():
    import run_inference_performance_test as perf

    out = perf.main(["--num-requests", "3", "--prompt-length", "12", "--num-tokens-to-generate", "4", "--inference-dynamic-batching-max-tokens", "8"])
    assert out["requests"] == 3 and out["throughput_tok_per_sec"] > 0 and out["prefill_chunks"] >= 3 * 2 - 1 and out["decode_forwards"] >= 3
