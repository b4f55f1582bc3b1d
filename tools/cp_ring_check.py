"""Context-parallel ring attention on N GPUs (NCCL p2p + the tcgen05 block kernels): zig-zag chunks of one long sequence per rank, forward + backward, checked
against plain causal attention over the full sequence computed on every rank.  torchrun --nproc-per-node N tools/cp_ring_check.py"""
import json, math, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from megatron_b200 import ops
    from megatron_b200.parallel.context_parallel import _AsyncRing, _RingAttnFn

    cp, c, hq, hk, d = world, int(os.environ.get("CHUNK", "2048")), 8, 2, 128
    s = 2 * cp * c
    torch.manual_seed(0)
    full = [torch.randn(s, 1, h, d, device="cuda").bfloat16() for h in (hq, hk, hk)]
    go_full = torch.randn(s, 1, hq, d, device="cuda").bfloat16()
    idx = torch.cat([torch.arange(rank * c, (rank + 1) * c), torch.arange((2 * cp - 1 - rank) * c, (2 * cp - rank) * c)]).cuda()
    q, k, v = (t[idx].clone().requires_grad_(True) for t in full)
    scale = 1 / math.sqrt(d)
    group = dist.group.WORLD
    shift = _AsyncRing(group)
    out = _RingAttnFn.apply(q, k, v, scale, True, rank, cp, shift)
    out.backward(go_full[idx])
    fq, fk, fv = (t.clone().requires_grad_(True) for t in full)
    want = ops.flash_attention(fq, fk, fv, causal=True, scale=scale)
    want.backward(go_full)
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()
    errs = {"out": rel(out, want[idx]), "dq": rel(q.grad, fq.grad[idx]), "dk": rel(k.grad, fk.grad[idx]), "dv": rel(v.grad, fv.grad[idx])}
    # timing: ring fwd+bwd vs the full-sequence kernels on one GPU (the ring does 1/cp of the pairs per rank)
    def fb_ring():
        o = _RingAttnFn.apply(q, k, v, scale, True, rank, cp, shift)
        o.backward(go_full[idx])
    def fb_full():
        o = ops.flash_attention(fq, fk, fv, causal=True, scale=scale)
        o.backward(go_full)
    times = {}
    for name, fn in (("ring", fb_ring), ("full_one_gpu", fb_full)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = round(t.item(), 3)
    worst = torch.tensor([max(errs.values())], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"bench": "cp_ring_attention", "cp": cp, "seq": s, "heads": f"{hq}/{hk}", "rel_err_rank0": {k_: round(v_, 5) for k_, v_ in errs.items()},
                          "max_rel_err_all_ranks": round(worst.item(), 5), "ms_fwd_bwd": times, "ok": bool(worst.item() < 3e-2)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
