"""Pad the number of tokens routed to every expert up to a multiple (reference ``fusions/fused_pad_routing_map.py``, a Triton kernel; FP8 grouped GEMMs need
16-aligned token counts).  Extra (token, expert) pairs are taken from tokens not yet routed to that expert, lowest index first.
CUDA: one block per expert, zeros ranked by a block-wide ballot scan (``ops/csrc/routing_kernels.cu``).  CPU: cumsum."""
import torch

from ... import ops


def fused_pad_routing_map(routing_map: torch.Tensor, pad_multiple: int) -> torch.Tensor:
    T, E = routing_map.shape
    if T == 0:
        return routing_map
    if routing_map.is_cuda and ops.has_ext() and hasattr(ops.ext(), "pad_routing_map"):
        out = ops.ext().pad_routing_map(routing_map.bool().contiguous(), int(pad_multiple))
        ops._count()
        return out
    rm = routing_map.bool()
    counts = rm.sum(0)
    need = (pad_multiple - counts % pad_multiple) % pad_multiple                     # [E]
    # rank of each unrouted token among the unrouted tokens of its expert column
    unrouted = ~rm
    rank = torch.cumsum(unrouted.to(torch.int64), dim=0) - 1
    add = unrouted & (rank < need.unsqueeze(0))
    return rm | add
