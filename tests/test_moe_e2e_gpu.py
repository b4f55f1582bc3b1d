"""End-to-end MoE training on 2 GPUs with expert parallelism: NVLink push/pull dispatcher + grouped tcgen05 GEMM vs the NCCL all-to-all
dispatcher — same losses, loss decreases (spawned ranks, one per GPU)."""
import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, port, q):
    import traceback

    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from megatron_b200 import ops
        from megatron_b200.core import parallel_state as ps
        from megatron_b200.parallel import collectives
        from megatron_b200.training.engine import TrainEngine

        losses = {}
        for kind in ("alltoall", "flex"):
            eng = TrainEngine("tiny_mixtral", expert_model_parallel_size=world, micro_batch_size=2, global_batch_size=4 * world, bf16=True, seed=1234,
                              model_overrides={"moe_token_dispatcher_type": kind})
            batch = eng.synthetic_batch(seed=7 + rank)
            ops.reset_launch_count()
            losses[kind] = [float(eng.train_step(batch)) for _ in range(4)]
            if kind == "flex":
                be = collectives.backend_for(ps.get_expert_model_parallel_group())
                assert be is not None, "flex dispatcher did not get an NVLink backend"
            del eng
            ps.destroy_model_parallel()
            torch.cuda.empty_cache()
        a, f = losses["alltoall"], losses["flex"]
        assert f[-1] < f[0], f"loss did not decrease: {f}"
        assert all(abs(x - y) < 0.05 for x, y in zip(a, f)), f"dispatchers disagree: {a} vs {f}"
        q.put((rank, "ok", losses))
    except Exception:
        q.put((rank, "err", traceback.format_exc()))


def test_moe_training_flex_dispatcher_matches_alltoall():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        for _ in range(2):
            rank, status, payload = q.get(timeout=300)
            assert status == "ok", f"rank {rank}:\n{payload}"
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
