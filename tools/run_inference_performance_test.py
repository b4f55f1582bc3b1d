#!/usr/bin/env python
"""Serving throughput / latency test of the dynamic engine (reference ``tools/run_inference_performance_test.py`` and
``tests/performance_tests/test_cases/gpt/gpt_16b_perf``): N requests of a given prompt / output length through continuous batching; reports generated
tokens/s, TPOT (ms per output token per request), TTFT and the engine's prefill / decode counters as one JSON line.

    python tools/run_inference_performance_test.py --preset llama3_8b --num-requests 32 --prompt-length 60 --num-tokens-to-generate 256 --enable-cuda-graphs
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny_llama")
    ap.add_argument("--num-layers", type=int, default=None)
    ap.add_argument("--num-requests", type=int, default=8)
    ap.add_argument("--prompt-length", type=int, default=60)
    ap.add_argument("--num-tokens-to-generate", type=int, default=32)
    ap.add_argument("--inference-dynamic-batching-block-size", type=int, default=16)
    ap.add_argument("--inference-dynamic-batching-max-requests", type=int, default=None)
    ap.add_argument("--inference-dynamic-batching-max-tokens", type=int, default=None, help="prompt tokens per step (chunked prefill)")
    ap.add_argument("--enable-cuda-graphs", action="store_true")
    ap.add_argument("--temperature", type=float, default=0.0)
    args = ap.parse_args(argv)
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29992")
    cuda = torch.cuda.is_available()
    if not dist.is_initialized():
        dist.init_process_group("nccl" if cuda else "gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model

    if not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    kw = dict(num_layers=args.num_layers) if args.num_layers else {}
    if cuda:
        kw.update(bf16=True, params_dtype=torch.bfloat16)
    else:
        kw.update(use_cpu_initialization=True)
    model, cfg, p = build_gpt_model(args.preset, **kw)
    model = (model.cuda() if cuda else model).eval()
    n, L, G = args.num_requests, args.prompt_length, args.num_tokens_to_generate
    max_running = args.inference_dynamic_batching_max_requests or n
    bs = args.inference_dynamic_batching_block_size
    eng = DynamicInferenceEngine(model, num_blocks=max_running * ((L + G) // bs + 2) + 16, block_size=bs, max_running=max_running, vocab_size=p["vocab_size"],
                                 enable_cuda_graphs=args.enable_cuda_graphs and cuda, max_prefill_tokens_per_step=args.inference_dynamic_batching_max_tokens)
    g = torch.Generator().manual_seed(0)
    ids = [eng.add_request(torch.randint(0, p["vocab_size"], (L,), generator=g).tolist(), SamplingParams(temperature=args.temperature, num_tokens_to_generate=G)) for _ in range(n)]
    sync = torch.cuda.synchronize if cuda else (lambda: None)
    sync()
    t0 = time.perf_counter()
    done = eng.run_until_done()
    sync()
    dt = time.perf_counter() - t0
    reqs = [done[i] for i in ids]
    gen = sum(len(r.generated_tokens) for r in reqs)
    ttft = [r.ttft for r in reqs if r.ttft is not None]
    tpot = [(r.finish_time - r.first_token_time) * 1e3 / max(len(r.generated_tokens) - 1, 1) for r in reqs if r.finish_time and r.first_token_time]
    out = {"bench": "inference_performance", "model": args.preset, "requests": n, "prompt_tokens": L, "output_tokens": G, "throughput_tok_per_sec": round(gen / dt, 1),
           "tpot_ms_mean": round(sum(tpot) / max(len(tpot), 1), 3), "ttft_ms_mean": round(1e3 * sum(ttft) / max(len(ttft), 1), 2), "steps": eng.steps,
           "decode_forwards": eng.decode_forwards, "prefill_chunks": eng.prefill_chunks, "cuda_graphs": len(eng._graphs), "device": "cuda" if cuda else "cpu"}
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    main()
