"""Serving benchmark on 1 GPU: (1) the paged-attention decode kernel alone (GB/s of K/V streamed) next to the round-1 formulation (block-table gather +
masked SDPA); (2) the dynamic engine end to end on a Llama-3-8B-shaped model: `BATCH` requests, ~60 prompt tokens, `OUT` generated tokens each —
throughput (generated tokens/s) and TPOT (ms per output token per request), the metrics of the reference's
tests/performance_tests/test_cases/gpt/gpt_16b_perf (H100: 324.8 tok/s, TPOT 98.5 ms at batch 32 for its 16B MoE).  Random-init weights, random prompts."""
import json, math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def kernel_bench():
    hq, hk, d, bs = 32, 8, 128, 16
    for B, L in [(32, 2048), (32, 512), (8, 8192), (128, 1024)]:
        width = (L + bs - 1) // bs
        nb = B * width + 1
        kp = torch.randn(nb, bs, hk, d, device="cuda").bfloat16(); vp = torch.randn_like(kp)
        table = torch.randperm(nb - 1, device="cuda")[: B * width].view(B, width).to(torch.int32)
        lengths = torch.full((B,), L, device="cuda", dtype=torch.int32)
        q = torch.randn(B, hq, d, device="cuda").bfloat16()
        scale = 1 / math.sqrt(d)
        t_k = timeit(lambda: ops.ext().paged_decode(q, kp, vp, table, lengths, scale, L))
        def gather_sdpa():
            K = kp[table.long()].reshape(B, width * bs, hk, d); V = vp[table.long()].reshape(B, width * bs, hk, d)
            rep = hq // hk
            qf = q.reshape(B * hk, rep, 1, d)
            Kf = K.permute(0, 2, 1, 3).reshape(B * hk, 1, -1, d).expand(B * hk, rep, width * bs, d)
            Vf = V.permute(0, 2, 1, 3).reshape(B * hk, 1, -1, d).expand(B * hk, rep, width * bs, d)
            return torch.nn.functional.scaled_dot_product_attention(qf, Kf, Vf, scale=scale)
        t_g = timeit(gather_sdpa, 10)
        gb = 2 * B * L * hk * d * 2 / 1e9
        print(json.dumps({"bench": "paged_decode_kernel", "B": B, "L": L, "heads": f"{hq}/{hk}", "ours_us": round(t_k * 1e3, 1), "ours_GBps": round(gb / t_k * 1e3, 0),
                          "gather_sdpa_us": round(t_g * 1e3, 1), "speedup": round(t_g / t_k, 2)}), flush=True)

def engine_bench():
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29991")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model
    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(1)
    layers = int(os.environ.get("LAYERS", "32"))
    m, cfg, p = build_gpt_model("llama3_8b", num_layers=layers, bf16=True, params_dtype=torch.bfloat16)
    m = m.cuda().eval()
    B, OUT, IN = int(os.environ.get("BATCH", "32")), int(os.environ.get("OUT", "256")), 60
    for batch in ([1, 8, B] if os.environ.get("SWEEP", "1") == "1" else [B]):
      for graphs in ([False, True] if os.environ.get("GRAPHS", "1") == "1" else [False]):
          eng = DynamicInferenceEngine(m, num_blocks=batch * ((IN + OUT) // 16 + 8) + 16, block_size=16, max_running=batch, vocab_size=p["vocab_size"], enable_cuda_graphs=graphs)
          g = torch.Generator().manual_seed(0)
          for _ in range(batch):
              eng.add_request(torch.randint(0, p["vocab_size"], (IN,), generator=g).tolist(), SamplingParams(temperature=0.0, num_tokens_to_generate=OUT))
          torch.cuda.synchronize(); t0 = time.perf_counter()
          eng.step()                                    # admission + prefill of every request (+ first token)
          torch.cuda.synchronize(); t1 = time.perf_counter()
          done = eng.run_until_done()
          torch.cuda.synchronize(); t2 = time.perf_counter()
          gen = sum(len(r.generated_tokens) for r in done.values())
          print(json.dumps({"bench": "dynamic_engine", "model": f"llama3_8b[{layers} layers]", "batch": batch, "prompt_tokens": IN, "output_tokens": OUT,
                            "throughput_tok_per_sec": round(gen / (t2 - t0), 1), "tpot_ms_per_tok": round((t2 - t1) * 1e3 / (OUT - 1), 2), "prefill_ms": round((t1 - t0) * 1e3, 1),
                            "cuda_graphs": graphs, "graphs_captured": len(eng._graphs), "decode_forwards": eng.decode_forwards, "first_tokens": list(done.values())[0].generated_tokens[:6], "reference_gpt16b_moe_H100": {"throughput": 324.8, "tpot_ms": 98.5, "batch": 32}}), flush=True)

if __name__ == "__main__":
    kernel_bench()
    if os.environ.get("ENGINE", "1") == "1":
        engine_bench()
