"""Serving extras on CPU: batched speculative verify semantics, the reference module layout of the inference package, the model inference wrapper."""
import pytest
import torch




def test_batched_speculative_verify_reference_semantics():
    """CPU formulation of the batched verify kernel: identical distributions accept everything and sample the bonus row; a greedy draft that disagrees with a
    one-hot target is cut at the first mismatch and corrected to the target's token; the follow-up token comes from the residual distribution."""
    import torch

    from megatron_b200.core.inference.speculative import verify_draft_tokens_batched

    torch.manual_seed(0)
    B, k, V = 6, 4, 50
    tp = torch.softmax(torch.randn(B, k + 1, V), -1)
    dt = torch.multinomial(tp[:, :k].reshape(-1, V), 1).view(B, k)
    n, nxt = verify_draft_tokens_batched(dt, tp[:, :k].clone(), tp)
    assert torch.equal(n, torch.full((B,), k)) and ((nxt >= 0) & (nxt < V)).all()
    target_tok = tp.argmax(-1)
    onehot = torch.nn.functional.one_hot(target_tok, V).float()
    draft = target_tok[:, :k].clone()
    draft[2, 1] = (draft[2, 1] + 1) % V
    draft[4, 0] = (draft[4, 0] + 1) % V
    n, nxt = verify_draft_tokens_batched(draft, None, onehot)
    assert n.tolist() == [4, 4, 1, 4, 0, 4]
    assert torch.equal(nxt, target_tok[torch.arange(B), n])
    # residual sampling never returns a token whose residual mass is zero
    dp = torch.softmax(torch.randn(B, k, V), -1)
    dt = torch.multinomial(dp.view(-1, V), 1).view(B, k)
    for _ in range(20):
        n, nxt = verify_draft_tokens_batched(dt, dp, tp)
        for b in range(B):
            if n[b] < k:
                assert tp[b, n[b], nxt[b]] > dp[b, n[b], nxt[b]]


def test_reference_module_layout_of_the_inference_package():
    """A script written against the reference imports these paths; they resolve to this framework's classes."""
    import importlib

    wanted = {
        "inference.sampling_params": ["SamplingParams"], "inference.common_inference_params": ["CommonInferenceParams"],
        "inference.inference_request": ["InferenceRequest", "DynamicInferenceRequest", "Status"], "inference.contexts": ["BaseInferenceContext", "StaticInferenceContext", "DynamicInferenceContext", "KVBlockAllocator"],
        "inference.contexts.static_context": ["StaticInferenceContext"], "inference.contexts.dynamic_context": ["DynamicInferenceContext", "ContextOverflowError", "TokenOverflowError"],
        "inference.engines": ["AbstractEngine", "StaticInferenceEngine", "DynamicInferenceEngine", "EngineSuspendedError"], "inference.engines.mcore_engine": ["MCoreEngine"],
        "inference.model_inference_wrappers.abstract_model_inference_wrapper": ["AbstractModelInferenceWrapper"],
        "inference.model_inference_wrappers.gpt.gpt_inference_wrapper": ["GPTInferenceWrapper"], "inference.model_inference_wrappers.inference_wrapper_config": ["InferenceWrapperConfig"],
        "inference.text_generation_controllers.text_generation_controller": ["TextGenerationController"], "inference.text_generation_server": ["MegatronServer"],
        "inference.apis": ["MegatronAsyncLLM", "SamplingParams"], "inference.headers": ["Headers"], "inference.data_parallel_inference_coordinator": ["DataParallelInferenceCoordinator"],
        "inference.async_stream": ["AsyncStream"], "inference.utils": ["Counter", "InferenceMode", "get_attention_mask"],
        "inference.communication_utils": ["broadcast_from_last_pipeline_stage", "broadcast_int_list", "broadcast_float_list", "send_to_next_pipeline_rank"],
        "models.hybrid": ["HybridModel", "HybridStack", "HybridStackSubmodules", "hybrid_stack_spec", "Symbols"], "models.hybrid.hybrid_layer_allocation": ["Symbols"],
        "models.gpt.experimental_attention_variant_module_specs": ["get_transformer_block_with_experimental_attention_variant_spec", "get_linear_attention_pattern"],
        "config": ["set_experimental_flag", "is_experimental_enabled"], "energy_monitor": ["EnergyMonitor"], "transformer.torch_layer_norm": ["WrappedTorchLayerNorm"],
        "distributed.torch_fully_sharded_data_parallel_config": ["TorchFullyShardedDataParallelConfig"],
    }
    for mod, names in wanted.items():
        m = importlib.import_module("megatron_b200.core." + mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"
    from megatron_b200.core import config
    from megatron_b200.core.inference.engines import AbstractEngine, DynamicInferenceEngine
    from megatron_b200.core.inference.inference_request import InferenceRequest, Status
    from megatron_b200.core.inference.utils import Counter
    from megatron_b200.core.utils import experimental_fn

    assert issubclass(DynamicInferenceEngine, AbstractEngine) and Status.of(InferenceRequest(0, [1])) is Status.WAITING_IN_QUEUE
    c = Counter()
    assert [next(c), next(c)] == [0, 1]
    c.reset()
    assert next(c) == 0

    @experimental_fn("0.1")
    def f():
        return 7

    config.set_experimental_flag(False)
    with pytest.raises(RuntimeError):
        f()
    config.set_experimental_flag(True)
    assert f() == 7
    config.set_experimental_flag(False)


def _wrapper_decode(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.contexts import StaticInferenceContext
    from megatron_b200.core.inference.model_inference_wrappers.gpt.gpt_inference_wrapper import GPTInferenceWrapper
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(3)
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0)
    m = GPTModel(cfg, get_gpt_layer_local_spec(), vocab_size=96, max_sequence_length=32, position_embedding_type="rope")
    w = GPTInferenceWrapper(m, StaticInferenceContext(2, 32))
    w.prep_model_for_inference()
    tok = torch.randint(0, 96, (2, 12), generator=torch.Generator().manual_seed(0))
    inp = w.prep_inference_input(tok)
    with torch.no_grad():
        full = m(tok, inp["position_ids"], None)
    # prefill 8 tokens, then four single-token decode steps against the KV cache in the context
    logits = [w.run_one_forward_step(w.get_batch_for_context_window(inp, 0, 8))]
    for t in range(8, 12):
        logits.append(w.run_one_forward_step(w.get_batch_for_context_window(inp, t, t + 1)))
    assert w.inference_context.sequence_len_offset == 12
    got = torch.cat(logits, dim=1)
    assert got.shape == full.shape and (got - full).abs().max().item() < 2e-4
    w.prep_model_for_inference()
    assert w.inference_context.sequence_len_offset == 0 and not w.inference_context.key_value_memory_dict
    return True


def test_gpt_inference_wrapper_incremental_decode_matches_full_forward():
    from dist_utils import run_distributed

    assert run_distributed(_wrapper_decode, 1) == [True]
