"""Pad the number of tokens routed to every expert up to a multiple (reference ``fusions/fused_pad_routing_map.py``; FP8 GEMMs need
16-aligned token counts).  Extra (token, expert) pairs are taken from tokens not yet routed to that expert, lowest index first."""
import torch


def fused_pad_routing_map(routing_map: torch.Tensor, pad_multiple: int) -> torch.Tensor:
    T, E = routing_map.shape
    rm = routing_map.bool()
    counts = rm.sum(0)
    need = (pad_multiple - counts % pad_multiple) % pad_multiple                     # [E]
    # rank of each unrouted token among the unrouted tokens of its expert column
    unrouted = ~rm
    rank = torch.cumsum(unrouted.to(torch.int64), dim=0) - 1
    add = unrouted & (rank < need.unsqueeze(0))
    return rm | add
