"""Spawn N CPU ranks (gloo) on this machine and run ``fn(rank, world, *args)`` in each."""
import os
import socket
import tempfile
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn, args, q, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        q.put((rank, "ok", out))
    except Exception:
        q.put((rank, "err", traceback.format_exc()))
    finally:
        try:
            from megatron_b200.core import parallel_state as ps

            ps.destroy_model_parallel()
            dist.destroy_process_group()
        except Exception:
            pass
        # tensors in the payload travel as file descriptors served by THIS process (torch.multiprocessing reductions): stay alive until the parent has
        # rebuilt them, otherwise it sees FileNotFoundError on the resource-sharer socket (a race that showed up under load)
        done.wait(120)


def run_distributed(fn, world, *args, timeout=300):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    done = ctx.Event()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q, done)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, payload = q.get(timeout=timeout)
            if status == "err":
                raise RuntimeError(f"rank {rank} failed:\n{payload}")
            results[rank] = payload
    finally:
        done.set()
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    return [results[r] for r in range(world)]
