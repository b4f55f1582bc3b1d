"""Rotary position embedding angles (reference ``models/common/embeddings/rotary_pos_embedding.py:36``)."""
from __future__ import annotations

import math
from typing import Optional

import torch

from megatron_b200 import ops
from megatron_b200.core import parallel_state as ps


def get_pos_emb_on_this_cp_rank(pos_emb: torch.Tensor, seq_dim: int, cp_group=None) -> torch.Tensor:
    """Pick the two zig-zag chunks of this CP rank (matches ``get_batch_on_this_cp_rank``)."""
    cp = ps.get_context_parallel_world_size() if cp_group is None else torch.distributed.get_world_size(cp_group)
    if cp <= 1:
        return pos_emb
    r = ps.get_context_parallel_rank() if cp_group is None else torch.distributed.get_rank(cp_group)
    idx = torch.tensor([r, 2 * cp - r - 1], device=pos_emb.device)
    shp = pos_emb.shape
    pe = pos_emb.view(*shp[:seq_dim], 2 * cp, -1, *shp[seq_dim + 1 :]).index_select(seq_dim, idx)
    return pe.view(*shp[:seq_dim], -1, *shp[seq_dim + 1 :])


class RotaryEmbedding(torch.nn.Module):
    def __init__(self, kv_channels: int, rotary_percent: float = 1.0, rotary_interleaved: bool = False,
                 seq_len_interpolation_factor: Optional[float] = None, rotary_base: int = 10000, rope_scaling: bool = False,
                 rope_scaling_factor: float = 8.0, use_cpu_initialization: bool = False, cp_group=None):
        super().__init__()
        dim = kv_channels if rotary_percent >= 1.0 else int(kv_channels * rotary_percent)
        self.rotary_interleaved = rotary_interleaved
        self.seq_len_interpolation_factor = seq_len_interpolation_factor
        dev = "cpu" if (use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        inv = 1.0 / (rotary_base ** (torch.arange(0, dim, 2, dtype=torch.float32, device=dev) / dim))
        if rope_scaling:
            inv = self._llama3_scaling(inv, rope_scaling_factor)
        self.inv_freq = inv
        self.cp_group = cp_group
        self._cache = {}

    @staticmethod
    def _llama3_scaling(inv_freq, factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192):
        """Llama-3.1 frequency-dependent NTK scaling."""
        low_wl = original_max_position_embeddings / low_freq_factor
        high_wl = original_max_position_embeddings / high_freq_factor
        wl = 2 * math.pi / inv_freq
        scaled = torch.where(wl > low_wl, inv_freq / factor, inv_freq)
        smooth = (original_max_position_embeddings / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
        smoothed = (1 - smooth) * inv_freq / factor + smooth * inv_freq
        mid = (wl >= high_wl) & (wl <= low_wl)
        return torch.where(mid, smoothed, scaled)

    def get_freqs_non_repeated(self, max_seq_len: int, offset: int = 0) -> torch.Tensor:
        seq = torch.arange(max_seq_len, device=self.inv_freq.device, dtype=self.inv_freq.dtype) + offset
        if self.seq_len_interpolation_factor is not None:
            seq = seq / self.seq_len_interpolation_factor
        return torch.outer(seq, self.inv_freq)

    @torch.no_grad()
    def forward(self, max_seq_len: int, offset: int = 0, packed_seq: bool = False, cp_group=None) -> torch.Tensor:
        """Angles ``[s, 1, 1, dim]`` (fp32).  Cached per (len, offset)."""
        key = (max_seq_len, offset, packed_seq)
        if key in self._cache:
            return self._cache[key]
        if self.inv_freq.device.type == "cpu" and torch.cuda.is_available() and not getattr(self, "_keep_cpu", False):
            pass
        freqs = self.get_freqs_non_repeated(max_seq_len, offset)
        if not self.rotary_interleaved:
            emb = torch.cat((freqs, freqs), dim=-1)
        else:
            emb = torch.stack((freqs.view(-1, 1), freqs.view(-1, 1)), dim=-1).view(freqs.shape[0], -1)
        emb = emb[:, None, None, :]
        cpg = cp_group or self.cp_group
        if not packed_seq:
            emb = get_pos_emb_on_this_cp_rank(emb, 0, cpg)
        self._cache[key] = emb
        return emb

    def get_rotary_seq_len(self, inference_context, transformer, transformer_input, transformer_config, packed_seq_params=None) -> int:
        if packed_seq_params is not None:
            return max(getattr(packed_seq_params, "max_seqlen_q", 0) or 0, getattr(packed_seq_params, "max_seqlen_kv", 0) or 0)
        if inference_context is not None:
            return inference_context.max_sequence_length
        if transformer is not None and transformer.input_tensor is not None:
            n = transformer.input_tensor.size(0)
        else:
            n = transformer_input.size(0)
        if transformer_config.sequence_parallel:
            n *= transformer_config.tensor_model_parallel_size
        return n * transformer_config.context_parallel_size

    def _apply(self, fn, *a, **k):
        self.inv_freq = fn(self.inv_freq)
        self._cache.clear()
        return super()._apply(fn, *a, **k)


def apply_rotary_pos_emb(t, freqs, config=None, cu_seqlens=None, mscale: float = 1.0, cp_group=None):
    """Functional entry (reference ``rope_utils.py:316``): bshd layout ``[s, b, h, d]``; THD when
    ``cu_seqlens`` is given (each packed sequence restarts at position 0)."""
    interleaved = bool(config.rotary_interleaved) if config is not None else False
    if cu_seqlens is None:
        return ops.apply_rope(t, freqs, interleaved, mscale)
    lens = (cu_seqlens[1:] - cu_seqlens[:-1]).tolist()
    outs = [ops.apply_rope(x.unsqueeze(1), freqs[: x.size(0)], interleaved, mscale).squeeze(1) for x in torch.split(t, lens)]
    return torch.cat(outs)
