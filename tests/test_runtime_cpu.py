"""Runtime services: resharding planner/executor, communication-memory shims, CUDA-graph manager fallback (CPU, gloo)."""
import torch

from dist_utils import run_distributed


def _reshard(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.dist_checkpointing.mapping import ShardedTensor
    from megatron_b200.core.resharding import reshard_state_dict

    full_w = torch.arange(8 * 6, dtype=torch.float32).view(8, 6)
    full_b = torch.arange(8, dtype=torch.float32)
    # source: sharded along dim 0 (column-parallel), bias replicated;  destination: sharded along dim 1 (row-parallel), bias sharded
    src = {"w": ShardedTensor.from_rank_offsets("w", full_w.chunk(world, 0)[rank].clone(), (0, rank, world)),
           "b": ShardedTensor.from_rank_offsets("b", full_b.clone())}
    dst = {"w": ShardedTensor.from_rank_offsets("w", torch.zeros(8, 6 // world), (1, rank, world)),
           "b": ShardedTensor.from_rank_offsets("b", torch.zeros(8 // world), (0, rank, world))}
    reshard_state_dict(src, dst, dist.group.WORLD)
    assert torch.equal(dst["w"].data, full_w.chunk(world, 1)[rank])
    assert torch.equal(dst["b"].data, full_b.chunk(world, 0)[rank])
    return True


def test_reshard_column_to_row_parallel():
    assert all(run_distributed(_reshard, 2))


def test_reshard_plan_prefers_local_replica_and_detects_holes():
    import pytest

    from megatron_b200.core.resharding import ShardDesc, build_reshard_plan

    src = [ShardDesc("w", (4, 4), (0, 0), (4, 4), 0), ShardDesc("w", (4, 4), (0, 0), (4, 4), 1)]
    dst = [ShardDesc("w", (4, 4), (0, 0), (2, 4), 0), ShardDesc("w", (4, 4), (2, 0), (2, 4), 1)]
    plan = build_reshard_plan(src, dst)
    assert all(op.src_rank == op.dst_rank for op in plan), "replicated sources: every rank should copy from itself"
    with pytest.raises(ValueError):
        build_reshard_plan([ShardDesc("w", (4, 4), (0, 0), (2, 4), 0)], [ShardDesc("w", (4, 4), (0, 0), (4, 4), 1)])


def test_cuda_graph_manager_falls_back_to_eager_on_cpu():
    from megatron_b200.core.transformer.cuda_graphs import graph_module

    net = graph_module(torch.nn.Linear(4, 4), warmup_steps=0)
    x = torch.randn(2, 4, requires_grad=True)
    net(x).sum().backward()
    assert x.grad is not None and net.cudagraph_manager.captured == {}


def test_nccl_mem_shim_allocates_plain_memory_without_gpu():
    from megatron_b200.core import nccl_allocator

    with nccl_allocator.nccl_mem(group=None) as mem:
        t = mem.alloc(16, torch.float32)
    assert t.shape == (16,) and float(t.abs().sum()) == 0.0


def _serve(rank, world):
    import asyncio
    import json
    import urllib.request

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.inference.text_generation import AsyncLLM, TextGenerationController, TextGenerationServer
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.tokenizers.tokenizer import build_tokenizer
    from megatron_b200.models.presets import build_gpt_model

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(7)
    model, _, p = build_gpt_model("tiny_llama", use_cpu_initialization=True)
    model.eval()
    tok = build_tokenizer("ByteLevel")
    greedy = SamplingParams(temperature=0.0, num_tokens_to_generate=6)
    ctl = TextGenerationController(DynamicInferenceEngine(model, vocab_size=256), tok)
    direct = ctl.generate(["hello", "b200"], greedy)
    srv = TextGenerationServer(ctl, port=0)
    srv.start()
    try:
        req = urllib.request.Request(f"http://127.0.0.1:{srv.port}/api", data=json.dumps({"prompts": ["hello", "b200"], "tokens_to_generate": 6, "temperature": 0.0}).encode(), method="PUT")
        out = json.loads(urllib.request.urlopen(req, timeout=60).read())
        req = urllib.request.Request(f"http://127.0.0.1:{srv.port}/v1/completions", data=json.dumps({"prompt": "hello", "max_tokens": 6, "temperature": 0.0}).encode(), method="POST")
        oai = json.loads(urllib.request.urlopen(req, timeout=60).read())
    finally:
        srv.stop()
    assert out["segments"] == [d["tokens"] for d in direct] and len(out["segments"][0]) == 6
    assert oai["choices"][0]["text"] == direct[0]["text"]

    async def many():
        llm = AsyncLLM(DynamicInferenceEngine(model, vocab_size=256), tok)
        return await asyncio.gather(*[llm.generate_tokens(list(tok.tokenize(s)), greedy) for s in ("hello", "b200", "hello")])

    reqs = asyncio.run(many())
    assert reqs[0].generated_tokens == reqs[2].generated_tokens == direct[0]["tokens"] and reqs[1].generated_tokens == direct[1]["tokens"]

    # streaming: token-list items per engine step, text deltas that concatenate to the blocking result, two streams interleaved
    async def streams():
        llm = AsyncLLM(DynamicInferenceEngine(model, vocab_size=256), tok)

        async def collect(s):
            return [d async for d in llm.generate_stream(s, greedy)]

        async def ids(s):
            return [i async for i in llm.stream_tokens(list(tok.tokenize(s)), greedy)]

        return await asyncio.gather(collect("hello"), collect("b200"), ids("hello"))

    a, b, c = asyncio.run(streams())
    assert "".join(a) == direct[0]["text"] and "".join(b) == direct[1]["text"]
    assert [t for item in c for t in item] == direct[0]["tokens"] and len(c) == 6               # one token per step

    # incremental detokenizer holds back an incomplete UTF-8 sequence
    from megatron_b200.core.inference.text_generation import IncrementalDetokenizer

    det = IncrementalDetokenizer(tok)
    euro = list("€".encode())                                                                   # 3 bytes
    assert det.add(list(b"a") + euro[:1]) == "a" and det.add(euro[1:2]) == "" and det.add(euro[2:]) == "€" and det.add(list(b"!"), final=True) == "!"

    # OpenAI chat + SSE streaming + health through the client
    from megatron_b200.core.inference.inference_client import InferenceClient

    srv = TextGenerationServer(TextGenerationController(DynamicInferenceEngine(model, vocab_size=256), tok), port=0)
    srv.start()
    try:
        cl = InferenceClient(port=srv.port)
        assert cl.health()["status"] == "ok"
        msgs = [{"role": "user", "content": "hi"}]
        chat = cl.chat(msgs, max_tokens=6, temperature=0.0)
        want = ctl.generate(["user: hi\nassistant: "], greedy)[0]
        assert chat["choices"][0]["message"]["content"] == want["text"] and chat["usage"]["completion_tokens"] == 6
        assert "".join(cl.stream(msgs, max_tokens=6, temperature=0.0)) == want["text"]
        assert "".join(cl.stream("hello", max_tokens=6, temperature=0.0)) == direct[0]["text"]
        assert cl.generate(["hello"], 6, temperature=0.0)[0] == "hello" + direct[0]["text"]
        try:
            cl.chat([{"role": "user"}])
            raise AssertionError("malformed message accepted")
        except urllib.error.HTTPError as e:
            assert e.code == 400
    finally:
        srv.stop()
    return True


def test_text_generation_server_and_async_llm():
    assert all(run_distributed(_serve, 1))
