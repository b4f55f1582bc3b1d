"""Serve a (randomly initialised or checkpointed) GPT preset over HTTP (reference ``tools/run_text_generation_server.py``).

    python tools/run_text_generation_server.py --preset tiny_llama --port 5000 [--load CKPT_DIR] [--engine dynamic]
    curl -X PUT localhost:5000/api -d '{"prompts": ["hello"], "tokens_to_generate": 16}'
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny_llama")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5000)
    ap.add_argument("--engine", choices=["static", "dynamic"], default="dynamic")
    ap.add_argument("--load", default=None)
    ap.add_argument("--tokenizer", default="byte")
    # the reference's engine flags (megatron/training/arguments.py _add_inference_args)
    ap.add_argument("--inference-dynamic-batching", action="store_true")
    ap.add_argument("--use-legacy-static-engine", action="store_true")
    ap.add_argument("--inference-dynamic-batching-block-size", type=int, default=None)
    ap.add_argument("--inference-dynamic-batching-max-requests", type=int, default=None)
    ap.add_argument("--inference-max-requests", type=int, default=None)
    ap.add_argument("--inference-dynamic-batching-max-tokens", type=int, default=None)
    ap.add_argument("--enable-chunked-prefill", action="store_true")
    ap.add_argument("--inference-dynamic-batching-prefix-caching", dest="inference_dynamic_batching_enable_prefix_caching", action="store_true")
    ap.add_argument("--inference-dynamic-batching-num-cuda-graphs", type=int, default=None)
    ap.add_argument("--decode-only-cuda-graphs", action="store_true")
    args = ap.parse_args()
    if args.use_legacy_static_engine:
        args.engine = "static"
    elif args.inference_dynamic_batching:
        args.engine = "dynamic"
    from megatron_b200.training.reference_flags import apply_reference_compat, engine_kwargs_from_args

    apply_reference_compat(args)
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29561")
    dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine, StaticInferenceEngine
    from megatron_b200.core.inference.text_generation import TextGenerationController, TextGenerationServer
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.tokenizers.tokenizer import MegatronTokenizer
    from megatron_b200.models.presets import build_gpt_model

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1234)
    model, _cfg, _p = build_gpt_model(args.preset)
    if args.load:
        from megatron_b200.core import dist_checkpointing

        sd = dist_checkpointing.load(model.sharded_state_dict(), args.load)
        model.load_state_dict(sd, strict=False)
    model.eval()
    from megatron_b200.core.tokenizers.tokenizer import build_tokenizer

    tok = build_tokenizer("ByteLevel") if args.tokenizer == "byte" else MegatronTokenizer.from_pretrained(args.tokenizer)
    eng = DynamicInferenceEngine(model, **engine_kwargs_from_args(args)) if args.engine == "dynamic" else StaticInferenceEngine(model, tok)
    srv = TextGenerationServer(TextGenerationController(eng, tok), args.host, args.port)
    print(f"serving {args.preset} on http://{args.host}:{args.port}/api", flush=True)
    srv.start(background=False)


if __name__ == "__main__":
    main()
