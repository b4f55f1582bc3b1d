"""InferenceConfig sizing, request serialisation and records (core/inference/{config,inference_request}.py)."""
import numpy as np
import pytest
import torch


class _MC:
    num_layers, num_attention_heads, num_query_groups, hidden_size, kv_channels, multi_latent_attention = 32, 32, 8, 4096, 128, False
    params_dtype = torch.bfloat16


def test_inference_config_sizes_kv_cache_and_graph_buckets():
    from megatron_b200.core.inference.config import CudaGraphSizingDistribution, InferenceConfig, MambaInferenceStateConfig

    cfg = InferenceConfig(buffer_size_gb=16, block_size_tokens=16, num_cuda_graphs=4, cuda_graph_sizing_distribution="exponential")
    per_tok = InferenceConfig.kv_bytes_per_token(_MC)
    assert per_tok == 2 * 32 * 8 * 128 * 2                                    # Llama-3-8B: 128 KiB of KV per token
    assert cfg.num_blocks(_MC) == (16 << 30) // (per_tok * 16)
    assert InferenceConfig.kv_bytes_per_token(_MC, tp_size=8) == per_tok // 8
    assert cfg.cuda_graph_batch_sizes(100) == [1, 16, 32, 64, 100]
    assert InferenceConfig(num_cuda_graphs=4, cuda_graph_sizing_distribution=CudaGraphSizingDistribution.LINEAR).cuda_graph_batch_sizes(64) == [16, 32, 48, 64]
    assert InferenceConfig(num_cuda_graphs=8, cuda_graph_sizing_distribution="mixed").cuda_graph_batch_sizes(64) == [1, 2, 4, 8, 16, 32, 48, 64]
    assert InferenceConfig().cuda_graph_batch_sizes(64) == []
    with pytest.raises(ValueError):
        InferenceConfig(block_size_tokens=24)
    m = MambaInferenceStateConfig(["M", "*", "M"], (8, 4), (2, 4, 16))
    assert m.bytes_per_request() == 2 * (8 * 4 * 2 + 2 * 4 * 16 * 4)
    hybrid = InferenceConfig(buffer_size_gb=1, mamba_inference_state_config=m, mamba_memory_ratio=0.25)
    assert hybrid.num_blocks(_MC) == int((1 << 30) * 0.75 // (per_tok * 16))

    class MLA(_MC):
        multi_latent_attention, kv_lora_rank, qk_pos_emb_head_dim = True, 512, 64
    assert InferenceConfig.kv_bytes_per_token(MLA, tp_size=8) == 32 * 576 * 2, "the latent cache is neither per-head nor TP-split"


def test_config_builds_a_working_engine():
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.config import InferenceConfig
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu,
                            add_bias_linear=False, normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=96, max_sequence_length=128, position_embedding_type="rope")
    icfg = InferenceConfig(buffer_size_gb=0.0005, block_size_tokens=8, max_requests=4, enable_prefix_caching=True, num_cuda_graphs=2)
    eng = icfg.build_engine(model, vocab_size=96)
    assert eng.cache.k.shape[1] == icfg.num_blocks(cfg, dtype=cfg.params_dtype) and eng.cache.block_size == 8
    rid = eng.add_request([1, 2, 3, 4, 5], SamplingParams(temperature=0.0, num_tokens_to_generate=4))
    assert len(eng.run_until_done()[rid].generated_tokens) == 4


def test_tensor_and_multimodal_serialisation_round_trip():
    from megatron_b200.core.inference import inference_request as R

    for t in (torch.randn(3, 4), torch.randn(5).bfloat16(), torch.arange(6, dtype=torch.int32).view(2, 3), torch.tensor([True, False])):
        back = R.deserialize_tensor(R.serialize_tensor(t))
        assert back.dtype == t.dtype and torch.equal(back, t)
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert np.array_equal(R.deserialize_ndarray(R.serialize_ndarray(a)), a)
    mm = {"images": [torch.rand(3, 8, 8), a], "image_sizes": [[8, 8], [3, 4]]}
    out = R.resolve_multimodal_data_for_engine(R.serialize_multimodal_data(mm), dtype=torch.bfloat16)
    assert out["images"][0].dtype == torch.bfloat16 and torch.allclose(out["images"][0].float(), mm["images"][0], atol=1e-2)
    assert torch.is_tensor(out["images"][1]) and out["image_sizes"] == [[8, 8], [3, 4]]


def test_block_hash_chains_events_and_records():
    from megatron_b200.core.inference import inference_request as R
    from megatron_b200.core.inference.engine import InferenceRequest
    from megatron_b200.core.inference.sampling import SamplingParams

    h = R.compute_block_hashes_batched([[1, 2, 3, 4, 5, 6, 7], [1, 2, 3, 4, 9, 9, 9, 9], [5, 6]], 4)
    assert len(h[0]) == 1 and len(h[1]) == 2 and h[2] == [] and h[0][0] == h[1][0] and h[1][1] != h[1][0]
    ev = R.DynamicInferenceEvent(R.DynamicInferenceEventType.GENERATED_TOKEN, payload=torch.tensor([7]))
    back = R.DynamicInferenceEvent.deserialize(ev.serialize())
    assert back.type is ev.type and torch.equal(back.payload, ev.payload)
    r1 = InferenceRequest(3, [1, 2, 3], SamplingParams(num_tokens_to_generate=8))
    r1.generated_tokens = [10, 11]
    rec = R.DynamicInferenceRequestRecord.from_request(r1)
    r2 = InferenceRequest(3, [1, 2, 3, 10, 11], SamplingParams(num_tokens_to_generate=6))      # re-admitted after a pause
    r2.generated_tokens = [12]
    rec.checkpoint(r2)
    rec.add_event(R.DynamicInferenceEventType.GENERATED_TOKEN)
    merged = rec.merge()
    assert merged.prompt_tokens == [1, 2, 3] and merged.generated_tokens == [10, 11, 12] and rec.time_to_first_token() is not None
    fin = R.FinishedRequestRecord(3, [1, 2, 3], [10, 11, 12], "abc", events=[ev])
    assert R.FinishedRequestRecord.deserialize(fin.serialize()).generated_tokens == [10, 11, 12]
    v = R.DynamicVLMInferenceRequest(0, [5, -200, 6, -200], [torch.zeros(1), torch.zeros(1)], num_img_embeddings=[4, 2])
    assert v.expanded_length() == 8 and v.image_spans() == [(1, 5), (6, 8)]
