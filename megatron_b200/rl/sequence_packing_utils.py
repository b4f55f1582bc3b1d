"""Sequence packing for RL batches (reference ``megatron/rl/sequence_packing_utils.py``, ``--rl-use-sequence-packing``).

Rollouts have very different lengths (prompt + however long the policy talked); padding them to the longest wastes most of the log-prob / training FLOPs.
Packing puts several rollouts back to back in one row of at most ``bin_size`` tokens and tells attention where the boundaries are (``PackedSeqParams`` with
``cu_seqlens`` — our attention kernels mask across boundaries natively, RoPE restarts per sequence), so the model sees ``[1, T]`` with no padding inside.

* ``pack_sequences``  — first-fit-decreasing bin packing (``algo='fifo'`` keeps arrival order instead, the reference's default for streaming consumption);
* ``PackedBatch``     — one bin: tokens / position ids / ``PackedSeqParams`` / per-token loss mask / which rollout every segment came from;
* ``packed_sequence_logprobs`` — per-token log-probs of the actual next tokens, un-packed back to one tensor per rollout."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from ..core.packed_seq_params import PackedSeqParams


def pack_sequences(lengths: Sequence[int], bin_size: int, algo: str = "ffd", max_sequences_per_bin: Optional[int] = None) -> List[List[int]]:
    """→ bins of sequence indices; every bin's total length ≤ ``bin_size`` (a sequence longer than the bin gets a bin of its own and must be truncated by the caller)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i]) if algo == "ffd" else list(range(len(lengths)))
    bins: List[List[int]] = []
    room: List[int] = []
    for i in order:
        n = lengths[i]
        placed = False
        candidates = range(len(bins)) if algo == "ffd" else range(max(len(bins) - 1, 0), len(bins))     # fifo only tries the open (last) bin
        for b in candidates:
            if room[b] >= n and (max_sequences_per_bin is None or len(bins[b]) < max_sequences_per_bin):
                bins[b].append(i)
                room[b] -= n
                placed = True
                break
        if not placed:
            bins.append([i])
            room.append(max(bin_size - n, 0))
    return bins


@dataclass
class PackedBatch:
    tokens: torch.Tensor              # [1, T]
    position_ids: torch.Tensor        # [1, T], restart at every boundary
    packed_seq_params: PackedSeqParams
    loss_mask: torch.Tensor           # [T - 1]: 1 where the TARGET (next token) is a completion token of the same sequence
    seq_index: List[int]              # original index of every packed sequence, in pack order
    lengths: List[int]


def build_packed_batch(seqs: Sequence[Sequence[int]], prompt_lens: Sequence[int], indices: Sequence[int], device="cpu") -> PackedBatch:
    toks, pos, mask, lens = [], [], [], []
    for i in indices:
        s, p = list(seqs[i]), prompt_lens[i]
        toks += s
        pos += list(range(len(s)))
        m = [0.0] * len(s)
        for t in range(max(p - 1, 0), len(s) - 1):          # position t predicts token t + 1; completion tokens are t + 1 ≥ p
            m[t] = 1.0
        mask += m                                            # the last position of a sequence predicts across the boundary: stays 0
        lens.append(len(s))
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0).tolist()), dtype=torch.int32, device=device)
    psp = PackedSeqParams(qkv_format="thd", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=max(lens), max_seqlen_kv=max(lens))
    return PackedBatch(torch.tensor([toks], device=device), torch.tensor([pos], device=device), psp, torch.tensor(mask[:-1], device=device), list(indices), lens)


def packed_sequence_logprobs(model, batch: PackedBatch, vocab_size: Optional[int] = None) -> torch.Tensor:
    """Log-prob of every actual next token of the pack, ``[T - 1]`` (entries that cross a boundary are meaningless and masked by ``loss_mask``)."""
    logits = model(batch.tokens, batch.position_ids, None, packed_seq_params=batch.packed_seq_params)
    if vocab_size is not None:
        logits = logits[..., :vocab_size]
    lp = torch.log_softmax(logits.float(), dim=-1)[0, :-1]
    return lp.gather(-1, batch.tokens[0, 1:].unsqueeze(-1)).squeeze(-1)


def unpack(values: torch.Tensor, batch: PackedBatch) -> List[torch.Tensor]:
    """[T - 1] per-token values → one ``[len_i - 1]`` tensor per packed sequence (the entry at each boundary is dropped)."""
    out, s0 = [], 0
    for n in batch.lengths:
        out.append(values[s0: s0 + n - 1])
        s0 += n
    return out
