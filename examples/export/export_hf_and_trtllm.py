#!/usr/bin/env python
"""Export a GPT preset to the Hugging Face Llama layout (and back), and write a TensorRT-LLM checkpoint directory.

    python examples/export/export_hf_and_trtllm.py --out /tmp/export_demo
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny_llama")
    ap.add_argument("--out", default="/tmp/export_demo")
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29578")
    dist.init_process_group("gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.export.hf_llama import hf_llama_to_megatron, megatron_to_hf_llama
    from megatron_b200.core.export.trtllm import ExportConfig, TRTLLMWeightsConverter, save_trtllm_checkpoint, trtllm_model_config
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    model, cfg, p = build_gpt_model(args.preset, use_cpu_initialization=True)
    sd = {k: v for k, v in model.state_dict().items() if isinstance(v, torch.Tensor)}
    dims = (cfg.num_attention_heads, cfg.num_query_groups, cfg.kv_channels)
    hf = megatron_to_hf_llama(sd, *dims)
    back = hf_llama_to_megatron(hf, *dims)
    worst = max((back[k].float() - v.float()).abs().max().item() for k, v in sd.items() if k in back)
    print(f"HF round trip: {len(hf)} tensors, max abs difference {worst:.2e}")
    os.makedirs(args.out, exist_ok=True)
    torch.save(hf, os.path.join(args.out, "hf_llama_state_dict.pt"))
    export = ExportConfig()
    weights = TRTLLMWeightsConverter(export, cfg).convert(sd, vocab_size=p["vocab_size"])
    config = trtllm_model_config(cfg, p["vocab_size"], p.get("seq_length", 2048), export)
    save_trtllm_checkpoint(os.path.join(args.out, "trtllm"), weights, config)
    print("TensorRT-LLM checkpoint:", sorted(os.listdir(os.path.join(args.out, "trtllm"))))


if __name__ == "__main__":
    main()
