"""Format conversion during a refit (reference ``resharding/transforms.py:23-295``).

The trainer holds bf16 weights; a serving model may hold them block-scaled (MXFP8: E4M3 payload + one E8M0 scale per 32
elements along K) in PERSISTENT buffers whose addresses are baked into the decode CUDA graphs.  A ``ReshardTransform`` lets the
receiver take the bf16 slice off the wire and write it, quantised, into the matching window of those buffers — the bf16 copy of
the weight never materialises at the destination."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch


class ReshardTransform:
    """Hooks around one transfer.  The default implementation is the identity on parameters it does not claim."""

    def should_transform(self, param_name: str) -> bool:
        return False

    def prepare_send(self, param_name: str, src_slice: Tuple[slice, ...], src_param: torch.Tensor) -> List[torch.Tensor]:
        """Tensors to put on the wire for this slice (default: the slice itself)."""
        return [src_param[src_slice]]

    def prepare_recv(self, param_name: str, dst_slice: Tuple[slice, ...]) -> List[torch.Tensor]:
        """Receive buffers (one per tensor ``prepare_send`` produces on the other side)."""
        raise NotImplementedError

    def finalize_recv(self, param_name: str, dst_slice: Tuple[slice, ...], recv_buffers: List[torch.Tensor]) -> None:
        """Called once the buffers are filled: write them into the destination storage."""
        raise NotImplementedError


def _scale_slice_from_data_slice(data_slice: Tuple[slice, ...], shape: Tuple[int, ...], block: int = 32) -> Tuple[slice, ...]:
    """Window of the [rows, K/32] scale tensor that belongs to a window of the [rows, K] payload.  The K window must start and
    end on block boundaries (TP splits of K are multiples of 32 for every supported model size)."""
    out = []
    for ax, (sl, n) in enumerate(zip(data_slice, shape)):
        lo, hi, _ = sl.indices(n)
        if ax == len(shape) - 1:
            if lo % block or hi % block:
                raise ValueError(f"MXFP8 refit: K window [{lo}, {hi}) is not aligned to the {block}-element scale blocks")
            out.append(slice(lo // block, hi // block))
        else:
            out.append(slice(lo, hi))
    return tuple(out)


def _ensure_sendable(param: torch.Tensor) -> torch.Tensor:
    """Plain dense tensor of a parameter that may be a quantised wrapper (``dequantize()``), on its own device."""
    deq = getattr(param, "dequantize", None)
    t = deq() if callable(deq) and type(param) not in (torch.Tensor, torch.nn.Parameter) else param
    return t.detach()


class MXFP8ReshardTransform(ReshardTransform):
    """Receive bf16, store MXFP8.

    ``buffers[name] = (payload uint8 [rows, K], scales uint8 [rows, K/32])`` are the serving model's persistent tensors
    (``core/post_training`` / ``inference`` quantised linears keep exactly this pair; the swizzled scale atoms the tcgen05 GEMM
    consumes are refreshed by ``refresh`` after the whole refit).  ``convertible`` limits the transform to the decoder GEMM
    weights; everything else (norms, embeddings, biases) takes the plain path."""

    def __init__(self, buffers: Dict[str, Tuple[torch.Tensor, torch.Tensor]], convertible: Optional[Iterable[str]] = None, refresh=None):
        self.buffers = buffers
        self.convertible = set(convertible) if convertible is not None else set(buffers)
        self.refresh = refresh
        self.touched: set = set()

    def should_transform(self, param_name: str) -> bool:
        return param_name in self.convertible and param_name in self.buffers

    def prepare_send(self, param_name, src_slice, src_param):
        return [_ensure_sendable(src_param)[src_slice].to(torch.bfloat16)]

    def prepare_recv(self, param_name, dst_slice):
        payload, _ = self.buffers[param_name]
        shape = tuple(len(range(*sl.indices(n))) for sl, n in zip(dst_slice, payload.shape))
        return [torch.empty(shape, dtype=torch.bfloat16, device=payload.device)]

    def finalize_recv(self, param_name, dst_slice, recv_buffers):
        from ... import ops
        payload, scales = self.buffers[param_name]
        x = recv_buffers[0]
        q, sf = ops.extra.mxfp8_quantize(x.reshape(-1, x.shape[-1]))
        payload[dst_slice].copy_(q.view(x.shape))
        s_sl = _scale_slice_from_data_slice(dst_slice, tuple(payload.shape))
        scales[s_sl].copy_(sf.view(scales[s_sl].shape))
        self.touched.add(param_name)

    def finish(self) -> None:
        """After the last transfer of a refit: let the owner rebuild derived layouts (swizzled scale atoms) once per weight."""
        if self.refresh is not None:
            for n in sorted(self.touched):
                self.refresh(n)
        self.touched.clear()
