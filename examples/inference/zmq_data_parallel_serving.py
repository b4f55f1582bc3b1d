#!/usr/bin/env python
"""Two data-parallel engine replicas behind the ZMQ coordinator, one client (single process, threads): routing by outstanding tokens, pause / resume, stats.

    python examples/inference/zmq_data_parallel_serving.py
"""
import os
import socket
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.inference.zmq_coordinator import ZMQInferenceClient, start_in_threads
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    engines = []
    for _ in range(2):
        torch.manual_seed(0)
        model, _, p = build_gpt_model("tiny_llama", use_cpu_initialization=not torch.cuda.is_available())
        engines.append(DynamicInferenceEngine(model.eval(), num_blocks=256, block_size=16, max_running=8, vocab_size=p["vocab_size"], max_prefill_tokens_per_step=64))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    coord, workers = start_in_threads(engines, port)
    client = ZMQInferenceClient(port)
    prompts = [[1 + (7 * i + j) % 50 for j in range(5 + 3 * i)] for i in range(6)]
    outs = client.generate(prompts, SamplingParams(temperature=0.0, num_tokens_to_generate=8))
    for p_, o in zip(prompts, outs):
        print(f"prompt[{len(p_):2d} tokens] -> {o}")
    print("stats:", client.stats())
    client.stop()
    coord.join(10)


if __name__ == "__main__":
    main()
