"""Collectives on quantised payloads (reference: FP8 / MXFP8 parameter all-gather ``fp8_param_gather``, TE's quantised sequence-parallel gather, the GTP
"quantised all-gathers"): the wire carries 1 byte/element + 1 scale byte per 32 elements instead of 2 bytes/element.

Row-wise MXFP8 quantisation commutes with gathering rows — every 1x32 block lives inside one row — so ``quantise → all-gather`` is bit-identical to
``all-gather → quantise`` at 53 % of the traffic, and the gathered (payload, scales) pair feeds the block-scaled GEMM directly: no dequantisation pass
exists anywhere on that path."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def all_gather_mxfp8(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``x [n, K]`` (this rank's rows) → (payload uint8 ``[world·n, K]``, scales uint8 ``[world·n, K/32]``), quantised BEFORE the wire."""
    from .. import ops

    q, sf = ops.mxfp8_quantize(x.to(torch.bfloat16))      # the CUDA quantiser consumes bf16; cast first so both paths round identically
    ws = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
    if ws == 1:
        return q, sf
    qg = q.new_empty((q.shape[0] * ws, q.shape[1]))
    sg = sf.new_empty((sf.shape[0] * ws, sf.shape[1]))
    dist.all_gather_into_tensor(qg, q.contiguous(), group=group)
    dist.all_gather_into_tensor(sg, sf.contiguous(), group=group)
    return qg, sg


def all_gather_dequantized(x: torch.Tensor, group=None, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Quantised all-gather with a dequantised result (parameter gather of ZeRO / GTP shards whose consumer wants bf16): half the bytes, ≈ 2^-4 relative
    rounding on the gathered copy only — the local fp32 / bf16 shard stays exact."""
    from .. import ops

    q, sf = all_gather_mxfp8(x.reshape(-1, x.shape[-1]), group)
    out = ops.mxfp8_dequantize(q, sf).to(dtype or x.dtype)
    ws = q.shape[0] // max(x.reshape(-1, x.shape[-1]).shape[0], 1)
    return out.view(x.shape[0] * ws, *x.shape[1:])


def wire_bytes(n_elements: int, quantised: bool) -> int:
    return n_elements + n_elements // 32 if quantised else 2 * n_elements
