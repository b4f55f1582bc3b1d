"""Batch-invariant (bitwise reproducible irrespective of batch size / co-scheduled requests) execution mode
(reference ``transformer/custom_layers/batch_invariant_kernels.py``: persistent Triton matmul, log-softmax, mean).

Sources of batch-size dependence and how this framework removes them when the mode is on:
* GEMM: different tile shapes / split-K per problem size change the reduction order.  Our tcgen05 GEMMs never split K; the mode pins ONE
  variant (1-CTA 128×256) instead of the per-shape autotuner choice, so every output element is always reduced in the same k-block order.
* log-softmax / mean / norms: reductions run per row with a fixed tree, independent of the number of rows — already invariant.
* attention: the KV block size (64) and the in-order accumulation over KV blocks do not depend on the batch — invariant; the mode forces
  the native kernel (library kernels may pick split-KV heuristics by batch size).
* paged decode: the flash-decoding kernel normally sizes its KV splits from (batch, longest request) to fill the GPU; the mode pins the split length to 512
  tokens, so split boundaries sit at fixed absolute positions and the partials of a request are combined in the same order whatever it is batched with
  (``tests/test_paged_attention_gpu.py::test_paged_decode_is_batch_invariant_in_the_mode``)."""
from __future__ import annotations

import contextlib
import os

import torch

_ENABLED = False
_SAVED = {}


def is_batch_invariant_mode_enabled() -> bool:
    return _ENABLED


def enable_batch_invariant_mode():
    global _ENABLED
    if _ENABLED:
        return
    from .... import ops
    from ....ops import gemm

    _SAVED["gemm"] = os.environ.get("MEGATRON_B200_GEMM")
    _SAVED["attn"] = ops._ATTN_IMPL
    os.environ["MEGATRON_B200_GEMM"] = "tcgen05:1"
    if hasattr(gemm, "set_mode"):
        gemm.set_mode("tcgen05:1")
    ops.set_attention_impl("native")
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
    _ENABLED = True


def disable_batch_invariant_mode():
    global _ENABLED
    if not _ENABLED:
        return
    from .... import ops
    from ....ops import gemm

    if _SAVED.get("gemm") is None:
        os.environ.pop("MEGATRON_B200_GEMM", None)
    else:
        os.environ["MEGATRON_B200_GEMM"] = _SAVED["gemm"]
    if hasattr(gemm, "set_mode"):
        gemm.set_mode(_SAVED.get("gemm") or "auto")
    ops.set_attention_impl(_SAVED.get("attn", "auto"))
    _ENABLED = False


@contextlib.contextmanager
def set_batch_invariant_mode(enabled: bool = True):
    was = _ENABLED
    (enable_batch_invariant_mode if enabled else disable_batch_invariant_mode)()
    try:
        yield
    finally:
        (enable_batch_invariant_mode if was else disable_batch_invariant_mode)()
