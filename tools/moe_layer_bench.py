"""One MoE layer, forward / backward milliseconds — the reference's ``tests/functional_tests/test_cases/common/moe_perf`` protocol
(``test_cases.py:60-68`` Mixtral proxy: seq 4096, micro-batch 1, hidden 4096, ffn 14336, 8 experts, top-2, EP = world, TP 1, bf16;
``__main__.py:33-37``: 5 warm-up + 20 timed iterations).  Reference goldens on 8×H100 (``baseline.json``): alltoall bf16 4.33 / 7.68 ms.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/moe_layer_bench.py
Times are CUDA events on the device, max over ranks.  One JSON line per dispatcher."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import torch.nn.functional as F

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.moe_module_specs import get_moe_module_spec
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.spec_utils import build_module
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.parallel import collectives

    ps.initialize_model_parallel(expert_model_parallel_size=world)
    model_parallel_cuda_manual_seed(1234)
    S, B, H, FFN, E, K = int(os.environ.get("SEQ", 4096)), 1, 4096, 14336, 8, 2
    for kind in os.environ.get("DISPATCHERS", "alltoall,flex").split(","):
        cfg = TransformerConfig(num_layers=1, hidden_size=H, ffn_hidden_size=FFN, num_attention_heads=32, num_query_groups=8, kv_channels=128, normalization="RMSNorm",
                                gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False, hidden_dropout=0.0, attention_dropout=0.0, bf16=True,
                                params_dtype=torch.bfloat16, expert_model_parallel_size=world, num_moe_experts=E, moe_router_topk=K, moe_token_dispatcher_type=kind,
                                moe_router_load_balancing_type="aux_loss", moe_aux_loss_coeff=1e-2, moe_grouped_gemm=True, gradient_accumulation_fusion=False)
        if kind == "flex":
            collectives.enable_for_group(ps.get_expert_model_parallel_group())
        layer = build_module(get_moe_module_spec(num_experts=E, moe_grouped_gemm=True), config=cfg).cuda()
        x = torch.randn(S, B, H, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        go = torch.randn(S, B, H, device="cuda", dtype=torch.bfloat16)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for it in range(25):
            dist.barrier()
            torch.cuda.synchronize()
            ev[0].record()
            y, _ = layer(x)
            ev[1].record()
            y.backward(go)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 5:
                tf += ev[0].elapsed_time(ev[1])
                tb += ev[1].elapsed_time(ev[2])
            x.grad = None
            for p in layer.parameters():
                p.grad = None
        t = torch.tensor([tf / 20, tb / 20], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"moe_layer": "mixtral proxy (seq 4096, mbs 1, h 4096, ffn 14336, 8 experts top-2)", "dispatcher": kind, "ep": world, "fwd_ms": round(float(t[0]), 3),
                              "bwd_ms": round(float(t[1]), 3), "reference_8xH100_alltoall_bf16_ms": [4.33, 7.68]}), flush=True)
        del layer
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
