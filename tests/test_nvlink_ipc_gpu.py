"""The cross-rank flag protocols of the NVLink collectives and of the fused GEMM<->collective kernels on ONE GPU.

Two (or four) processes share ``cuda:0``; each allocates its "symmetric heap" with the normal allocator and maps the
peers' heaps through CUDA IPC (``NVLinkBackend(peer_heaps=...)``), so every kernel runs its P2P (non-multicast)
protocol — peer loads/stores, ``st.release.sys`` / ``ld.acquire.sys`` chunk flags, epoch barriers — between real,
independently scheduled processes.  The control plane is gloo.  This is the correctness proof a 1-GPU CI box can give
for the multi-GPU kernels (the NVLS ``multimem`` variants need >= 2 GPUs: ``tests/test_nvlink_gpu.py``).

Covers: back-to-back reuse of the double-buffered workspaces (``repeats``), one deliberately late rank
(``skew_rank``), multi-chunk shards (chunks_per_rank = 4), all eight SP pair ops + the two all-reduce pair ops,
forward and backward, against gloo collectives + fp32 matmuls.
"""
import os
import socket
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu]

HEAP_BYTES = (1 << 16) + 2 * (96 << 20)


def _worker(rank, world, port, queues, q, case):
    import traceback

    try:
        os.environ.pop("PYTORCH_CUDA_ALLOC_CONF", None)       # cudaIpc handles need plain cudaMalloc segments
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
        import torch.distributed as dist

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from megatron_b200.parallel import collectives, fused
        from megatron_b200.parallel.nvlink import NVLinkBackend
        from megatron_b200.parallel.selfcheck import pair_op_self_check

        heap = torch.zeros(HEAP_BYTES, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for d in range(world):
            if d != rank:
                queues[d].put((rank, heap))
        peers = {rank: heap}
        for _ in range(world - 1):
            r, t = queues[rank].get(timeout=120)
            peers[r] = t
        g = dist.group.WORLD
        be = NVLinkBackend(g, peer_heaps=[peers[r] for r in range(world)])
        collectives._BACKENDS[id(g)] = be
        out = {}
        if case == "collectives":
            torch.manual_seed(rank)
            for it in range(3):
                x = torch.randn(512, 1024, device="cuda").bfloat16()
                xs = [torch.empty_like(x).cpu() for _ in range(world)]
                dist.all_gather(xs, x.cpu())
                got = be.all_gather(x)
                assert torch.equal(got.cpu(), torch.cat(xs)), f"all_gather mismatch iter {it}"
                y = torch.randn(world * 256, 512, device="cuda").bfloat16()
                ref = y.float().cpu()
                dist.all_reduce(ref)
                got = be.reduce_scatter(y)
                mine = ref[rank * 256 : (rank + 1) * 256]
                assert torch.allclose(got.float().cpu(), mine, atol=0.06, rtol=0.02), f"reduce_scatter mismatch iter {it}"
                z = torch.randn(777 * 8, device="cuda")
                ref = z.cpu().clone()
                dist.all_reduce(ref)
                got = be.all_reduce(z.clone())
                assert torch.allclose(got.cpu(), ref, atol=1e-4, rtol=1e-4), "all_reduce fp32 mismatch"
            out["ok"] = True
        else:
            fused.set_mode("fused")
            res = pair_op_self_check(g, seq=world * 1024, hidden=1024, ffn=2048, qkv=1536, quick=False, repeats=3, skew_rank=world - 1)
            out = res
            out["fused_calls"] = be.fused_calls
        torch.cuda.synchronize()
        dist.barrier()
        q.put((rank, "ok", out))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "err", traceback.format_exc()))


def _run(world, case, timeout=420):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    queues = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_worker, args=(r, world, port, queues, q, case)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=timeout) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    for rank, status, payload in res:
        assert status == "ok", f"rank {rank}:\n{payload}"
    return [r[2] for r in sorted(res)]


def test_ipc_collectives_two_ranks_one_gpu():
    _run(2, "collectives")


@pytest.mark.parametrize("world", [2, 4])
def test_ipc_fused_pair_ops_one_gpu(world):
    res = _run(world, "fused")
    r0 = res[0]
    print({k: v for k, v in r0.items()})
    assert r0["mode"] == "fused"
    assert r0["fused_calls"] >= 3 * 10, "the fused kernels did not run"
    assert r0["max"] < 2.5e-2, r0
