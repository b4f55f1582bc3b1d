"""Fused NVLink MoE dispatch/combine vs a gather-everything reference, forward and backward (>= 2 GPUs, spawned ranks)."""
import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, port, q):
    import traceback

    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        import torch.distributed as dist

        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from megatron_b200.parallel import collectives

        be = collectives.enable_for_group(dist.group.WORLD)
        T, H, L, K = 96, 256, 2, 2
        E = L * world
        for it in range(3):
            torch.manual_seed(100 * it + rank)
            tokens = torch.randn(T, H, device="cuda").bfloat16().requires_grad_(True)
            ids = torch.stack([torch.randperm(E, device="cuda")[:K] for _ in range(T)])
            rmap = torch.zeros(T, E, dtype=torch.bool, device="cuda").scatter(1, ids, True)
            probs = (torch.rand(T, E, device="cuda") * rmap).requires_grad_(True)
            h, recv, recv_probs, tpe = be.moe_dispatch(tokens, rmap, probs, L, topk=K)
            # reference: everybody sees everything
            all_tok = [torch.empty_like(tokens) for _ in range(world)]
            all_map = [torch.empty_like(rmap) for _ in range(world)]
            all_pr = [torch.empty_like(probs) for _ in range(world)]
            dist.all_gather(all_tok, tokens.detach())
            dist.all_gather(all_map, rmap)
            dist.all_gather(all_pr, probs.detach())
            exp_rows, exp_probs, exp_tpe = [], [], []
            for le in range(L):
                e = rank * L + le
                n = 0
                for src in range(world):
                    sel = all_map[src][:, e]
                    exp_rows.append(all_tok[src][sel])
                    exp_probs.append(all_pr[src][sel, e])
                    n += int(sel.sum())
                exp_tpe.append(n)
            assert tpe.tolist() == exp_tpe, (tpe.tolist(), exp_tpe)
            assert torch.equal(recv, torch.cat(exp_rows)), f"dispatch rows differ (iter {it})"
            assert torch.allclose(recv_probs, torch.cat(exp_probs)), "dispatched probs differ"
            # experts: y = 2 * x * prob  → combine: token t gets 2 * x_t * Σ_k p_tk
            y = recv * 2.0 * recv_probs.unsqueeze(-1).to(recv.dtype)
            out = be.moe_combine(y, h)
            ref = 2.0 * tokens.detach().float() * (probs.detach() * rmap).sum(1, keepdim=True)
            assert torch.allclose(out.float(), ref, atol=0.08, rtol=0.05), f"combine mismatch {(out.float() - ref).abs().max()}"
            out.float().sum().backward()
            gref = 2.0 * (probs.detach() * rmap).sum(1, keepdim=True).expand(-1, H)
            assert torch.allclose(tokens.grad.float(), gref, atol=0.08, rtol=0.05), "token grads through dispatch+combine"
            pg_ref = 2.0 * tokens.detach().float().sum(1, keepdim=True) * rmap
            assert torch.allclose(probs.grad, pg_ref, atol=0.5, rtol=0.05), f"prob grads {(probs.grad - pg_ref).abs().max()}"
        torch.cuda.synchronize()
        dist.barrier()
        q.put((rank, "ok", None))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "err", traceback.format_exc()))


def test_nvlink_moe_dispatch_combine():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 8)
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        for _ in range(world):
            rank, status, payload = q.get(timeout=240)
            assert status == "ok", f"rank {rank}:\n{payload}"
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
