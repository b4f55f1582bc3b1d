"""Program initialisation (reference ``training/initialize.py``): parse args → global vars → distributed + model-parallel groups → seeds →
native extensions.  ``initialize_megatron`` itself lives in ``training.py`` (kept there with the loop it serves); this module holds the pieces
and the reference's entry names."""
from __future__ import annotations

import os
import random
import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ..core import parallel_state as ps
from ..core.tensor_parallel.random import model_parallel_cuda_manual_seed
from .training import initialize_megatron  # noqa: F401  (re-export)


def set_random_seed(seed: int, data_parallel_random_init: bool = False, te_rng_tracker: bool = False, inference_rng_tracker: bool = False) -> int:
    """Seed python / numpy / torch; pipeline stages get different seeds (``+100·pp_rank``), DP replicas only when asked to."""
    if seed is None or seed <= 0:
        raise ValueError(f"seed ({seed}) should be a positive integer")
    if ps.is_initialized():
        seed = seed + 100 * ps.get_pipeline_model_parallel_rank()
        if data_parallel_random_init:
            seed = seed + 10 * ps.get_data_parallel_rank()
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if ps.is_initialized():
        model_parallel_cuda_manual_seed(seed)
    return seed


_set_random_seed = set_random_seed


def compile_dependencies(verbose: bool = False) -> None:
    """Build the native pieces once per node: rank 0 compiles (dataset index helpers, CUDA ops), the others wait at a barrier —
    the in-tree ``.so`` files are then visible to everyone through the shared filesystem."""
    first = not dist.is_initialized() or dist.get_rank() == 0
    t0 = time.time()
    if first:
        from ..ops import build as ops_build

        ops_build.build_datasets_helpers()
        if torch.cuda.is_available() or os.environ.get("MEGATRON_B200_BUILD_OPS", "0") == "1":
            ops_build.build_all()
    if dist.is_initialized():
        dist.barrier()
    if verbose and first:
        print(f"> native dependencies ready in {time.time() - t0:.1f}s", flush=True)


_compile_dependencies = compile_dependencies


def set_jit_fusion_options() -> None:
    """The reference configures the TorchScript / nvFuser fuser here; the fused ops of this framework are hand-written kernels, so the only
    thing left to set is that PyTorch does not try to fuse around them."""
    try:
        torch._C._jit_set_profiling_executor(False)
        torch._C._jit_set_profiling_mode(False)
    except Exception:
        pass


def init_autoresume(args=None) -> Optional[object]:
    """Cluster auto-resume hook (ADLR-internal in the reference): honour a ``MEGATRON_B200_AUTORESUME_FILE`` sentinel — when the file
    appears the loop checkpoints and exits with code 0 so the scheduler requeues the job."""
    path = os.environ.get("MEGATRON_B200_AUTORESUME_FILE")
    if not path:
        return None

    class _AutoResume:
        def termination_requested(self) -> bool:
            return os.path.exists(path)

        def request_resume(self) -> None:
            try:
                os.remove(path)
            except OSError:
                pass

    return _AutoResume()


def setup_nccl_flight_recorder(dump_dir: Optional[str] = None, buffer_size: int = 2000, dump_on_timeout: bool = True) -> dict:
    """Wire PyTorch's NCCL flight recorder for hang / desync post-mortems (reference ``initialize.py:291-333``): a ring of the last ``buffer_size`` collectives per
    rank, dumped to ``<dump_dir>/nccl_trace_rank_<r>`` when the watchdog fires.  Must run BEFORE ``init_process_group``."""
    env = {"TORCH_NCCL_TRACE_BUFFER_SIZE": str(buffer_size), "TORCH_NCCL_DUMP_ON_TIMEOUT": "1" if dump_on_timeout else "0", "TORCH_NCCL_ASYNC_ERROR_HANDLING": "1"}
    if dump_dir:
        os.makedirs(dump_dir, exist_ok=True)
        env["TORCH_NCCL_DEBUG_INFO_TEMP_FILE"] = os.path.join(dump_dir, "nccl_trace_rank_")
    for k, v in env.items():
        os.environ.setdefault(k, v)
    return env


def update_pg_timeout(timeout_minutes: float, groups=None) -> int:
    """Shorten / lengthen the collective timeout of live process groups (reference ``parallel_state.update_pg_timeout``): long during start-up and
    checkpoint loading, short in steady state so a dead rank is noticed quickly.  Returns how many groups were updated."""
    from datetime import timedelta

    if not dist.is_initialized():
        return 0
    setter = getattr(dist.distributed_c10d, "_set_pg_timeout", None)
    if setter is None:
        return 0
    n = 0
    for g in (groups if groups is not None else [dist.group.WORLD]):
        try:
            setter(timedelta(minutes=timeout_minutes), g)
            n += 1
        except Exception:       # gloo groups have no adjustable watchdog
            pass
    return n
