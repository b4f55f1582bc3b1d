"""Deterministic-execution mode and its self-check (reference ``training/determinism.py`` + the ``--deterministic-mode`` plumbing in
``arguments.py``): what has to be pinned on B200 so two runs of the same job are bit-identical.

* cuBLAS workspace config + ``torch.use_deterministic_algorithms`` for library ops that remain;
* our GEMM autotuner is frozen to one variant and attention to the native kernel (``batch_invariant_kernels``), because "fastest of N" is
  timing-dependent and split-K / stream-K style reductions reorder sums;
* NVLink reductions: ``multimem.ld_reduce`` order is fixed by the switch for a given topology; the NCCL fallback is pinned to the ring algorithm;
* atomics: the native attention backward accumulates dQ with ``red.add`` (order-dependent) → deterministic mode keeps the library backward.
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict, Iterable, List

import torch


def enable_deterministic_mode(strict: bool = True) -> Dict[str, str]:
    env = {"CUBLAS_WORKSPACE_CONFIG": ":4096:8", "NCCL_ALGO": "Ring", "NVTE_ALLOW_NONDETERMINISTIC_ALGO": "0",
           "MEGATRON_B200_ATTN_BWD": "library"}
    for k, v in env.items():
        os.environ[k] = v
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.use_deterministic_algorithms(True, warn_only=not strict)
    try:
        from ..core.transformer.custom_layers.batch_invariant_kernels import enable_batch_invariant_mode

        enable_batch_invariant_mode()
    except Exception:
        pass
    return env


def disable_deterministic_mode() -> None:
    torch.use_deterministic_algorithms(False)
    torch.backends.cudnn.deterministic = False
    try:
        from ..core.transformer.custom_layers.batch_invariant_kernels import disable_batch_invariant_mode

        disable_batch_invariant_mode()
    except Exception:
        pass


def tensor_digest(tensors: Iterable[torch.Tensor]) -> str:
    """Order-sensitive SHA-256 over the raw bytes of the tensors (bit-exactness check, not a numerical tolerance)."""
    h = hashlib.sha256()
    for t in tensors:
        c = t.detach().contiguous().cpu()
        h.update(str(c.dtype).encode())
        h.update(str(tuple(c.shape)).encode())
        h.update(c.view(torch.uint8).numpy().tobytes() if c.numel() else b"")
    return h.hexdigest()


def model_digest(model) -> str:
    models = model if isinstance(model, (list, tuple)) else [model]
    return tensor_digest(p for m in models for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]))


def check_determinism(run_fn, repeats: int = 2) -> List[str]:
    """Run ``run_fn() -> iterable of tensors`` ``repeats`` times; raises if the digests differ, returns them otherwise."""
    digests = [tensor_digest(run_fn()) for _ in range(repeats)]
    if len(set(digests)) != 1:
        raise RuntimeError(f"non-deterministic results: {digests}")
    return digests
