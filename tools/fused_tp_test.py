"""Numerics + timing of the TP pair ops (AG→GEMM, GEMM→RS, both backwards) in nccl / nvlink / fused modes.

Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/fused_tp_test.py
Times are CUDA-event times on the device, max over ranks.  Llama-3 8B shapes (seq 8192, mbs 1, hidden 4096, ffn 14336).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


_LOCAL = [False]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from megatron_b200 import ops
    from megatron_b200.parallel import collectives, fused

    g = dist.group.WORLD
    be = collectives.enable_for_group(g)
    S = int(os.environ.get("SEQ", "8192"))
    H, FFN, QKV = 4096, 14336, 6144
    iters = int(os.environ.get("ITERS", "20"))
    torch.manual_seed(1234 + rank)
    dev = "cuda"

    def rnd(*shape, scale=1.0):
        return (scale * torch.randn(*shape, device=dev)).bfloat16()

    cases = {
        "col_fwd_qkv": ("ag_gemm", rnd(S // world, 1, H), rnd(QKV // world, H, scale=0.02)),
        "col_fwd_fc1": ("ag_gemm", rnd(S // world, 1, H), rnd(2 * FFN // world, H, scale=0.02)),
        "row_fwd_proj": ("gemm_rs", rnd(S, 1, H // world), rnd(H, H // world, scale=0.02)),
        "row_fwd_fc2": ("gemm_rs", rnd(S, 1, FFN // world), rnd(H, FFN // world, scale=0.02)),
        "col_bwd_fc1": ("col_bwd", rnd(S, 1, 2 * FFN // world), rnd(S // world, 1, H), rnd(2 * FFN // world, H, scale=0.02)),
        "col_bwd_qkv": ("col_bwd", rnd(S, 1, QKV // world), rnd(S // world, 1, H), rnd(QKV // world, H, scale=0.02)),
        "row_bwd_fc2": ("row_bwd", rnd(S // world, 1, H), rnd(S, 1, FFN // world), rnd(H, FFN // world, scale=0.02)),
        "row_bwd_proj": ("row_bwd", rnd(S // world, 1, H), rnd(S, 1, H // world), rnd(H, H // world, scale=0.02)),
    }

    def run_local(kind, args):
        """The GEMMs of the pair op alone on full-size local operands (no communication): the compute floor."""
        if kind == "ag_gemm":
            return (ops.gemm_nt(local_full[id(args[0])], args[1]),)
        if kind == "gemm_rs":
            return (ops.gemm_nt(args[0], args[1]),)
        if kind == "col_bwd":   # dgrad + wgrad
            return ops.gemm_nn(args[0], args[2]), ops.gemm_tn(args[0].reshape(-1, args[0].shape[-1]), local_full[id(args[1])].reshape(-1, args[1].shape[-1]), out_dtype=args[2].dtype)
        full_gy = local_full[id(args[0])]
        return ops.gemm_nn(full_gy, args[2]), ops.gemm_tn(full_gy.reshape(-1, full_gy.shape[-1]), args[1].reshape(-1, args[1].shape[-1]), out_dtype=args[2].dtype)

    local_full = {}
    for name, (kind, *args) in cases.items():
        for t in args:
            if t.dim() == 3 and t.shape[0] == S // world:
                local_full[id(t)] = t.repeat(world, 1, 1)

    def run(kind, args):
        if _LOCAL[0]:
            return run_local(kind, args)
        if kind == "ag_gemm":
            return (fused.all_gather_gemm(args[0], args[1], g),)
        if kind == "gemm_rs":
            return (fused.gemm_reduce_scatter(args[0], args[1], g),)
        if kind == "col_bwd":
            return fused.sp_linear_backward(args[0], args[1], args[2], g, True, False)
        return fused.row_linear_backward_sp(args[0], args[1], args[2], g, True, False)

    _LOCAL[0] = False
    results = {}
    refs = {}
    modes = os.environ.get("MODES", "nccl,nvlink,fused:4:8,fused:8:12,fused:12:20").split(",")
    for mode in modes:
        _LOCAL[0] = mode == "local"      # "local" = GEMMs only, no communication (compute floor; outputs are not compared)
        if mode == "local":
            fused.set_mode("nccl")
        elif mode.startswith("fused:"):  # fused:<AG comm clusters>:<RS comm clusters>
            _, a, r = mode.split(":")
            be.fused_comm_clusters = [int(a), int(r)]
            fused.set_mode("fused")
        else:
            fused.set_mode(mode)
        for name, (kind, *args) in cases.items():
            outs = run(kind, args)
            outs = [o.float().clone() for o in outs]
            torch.cuda.synchronize()
            if mode == modes[0]:
                refs[name] = outs
            elif mode != "local":
                for i, (o, r) in enumerate(zip(outs, refs[name])):
                    err = (o - r).abs().max().item()
                    tol = 0.02 * r.abs().max().item() + 1e-2
                    if not (err <= tol):
                        raise AssertionError(f"[rank {rank}] {mode}:{name} output {i} max err {err} > {tol}")
            # timing
            for _ in range(3):
                run(kind, args)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run(kind, args)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            results.setdefault(name, {})[mode] = round(t.item() * 1e3, 1)
        if rank == 0:
            print(f"mode {mode} done (fused kernel launches so far: {be.fused_calls})", flush=True)
    fused.set_mode("nccl")
    if rank == 0 and "local" in modes:
        # exposed (non-overlapped) communication per op = time(mode) - time(GEMMs alone); per step = sum over the 8 pair ops x layers x micro-batches
        layers, mb = 32, 4
        for m in modes:
            if m != "local":
                per_layer = sum(max(0.0, results[n][m] - results[n]["local"]) for n in results)
                print(f"exposed communication [{m}]: {per_layer:.0f} us/layer/micro-batch -> {per_layer * layers * mb / 1e3:.1f} ms/step (32 layers x 4 micro-batches)", flush=True)
    if rank == 0:
        print(json.dumps({"world": world, "seq": S, "unit": "us", "multicast": bool(be.mc), "results": results}, indent=1), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
