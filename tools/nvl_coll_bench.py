"""Bandwidth of our NVLink collectives (multimem / P2P kernels) vs NCCL, swept over CTA count and message size.

Run under torchrun (see tools/fused_tp_test.py).  Device-side CUDA-event timing, max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from megatron_b200.parallel import collectives

    g = dist.group.WORLD
    be = collectives.enable_for_group(g)
    rows = []
    for mib in (8, 64):  # full (gathered / pre-scatter) message size
        n = mib * (1 << 20) // 2
        shard = torch.randn(n // world, device="cuda").bfloat16()
        full = torch.randn(n, device="cuda").bfloat16()
        out_full = torch.empty(n, device="cuda", dtype=torch.bfloat16)
        out_shard = torch.empty(n // world, device="cuda", dtype=torch.bfloat16)
        row = {"MiB": mib, "nccl_ag": timeit(lambda: dist.all_gather_into_tensor(out_full, shard)),
               "nccl_rs": timeit(lambda: dist.reduce_scatter_tensor(out_shard, full)), "nccl_ar": timeit(lambda: dist.all_reduce(full))}
        ws = be.symmetric_like((world, n // world), torch.bfloat16)
        ws.copy_(full.view(world, -1))
        for nb in (16, 32, 64, 128):
            be.nblocks = nb
            row[f"own_ag_{nb}"] = timeit(lambda: be.all_gather(shard))
            row[f"own_rs_{nb}"] = timeit(lambda: be.reduce_scatter(ws))
        rows.append({k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items()})
    if rank == 0:
        print(json.dumps({"world": world, "unit": "us", "multicast": bool(be.mc), "rows": rows}, indent=1), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
