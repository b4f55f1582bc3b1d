"""YaRN rotary embedding (reference ``yarn_rotary_pos_embedding.py``)."""
from __future__ import annotations

import math

import torch

from .rotary_pos_embedding import RotaryEmbedding


def _find_correction_dim(num_rot, dim, base, max_pos):
    return (dim * math.log(max_pos / (num_rot * 2 * math.pi))) / (2 * math.log(base))


def _find_correction_range(low_rot, high_rot, dim, base, max_pos):
    low = math.floor(_find_correction_dim(low_rot, dim, base, max_pos))
    high = math.ceil(_find_correction_dim(high_rot, dim, base, max_pos))
    return max(low, 0), min(high, dim - 1)


def _linear_ramp_mask(lo, hi, dim, device):
    if lo == hi:
        hi += 0.001
    return torch.clamp((torch.arange(dim, dtype=torch.float32, device=device) - lo) / (hi - lo), 0, 1)


def yarn_get_mscale(scale: float = 1.0, mscale: float = 1.0) -> float:
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


class YarnRotaryEmbedding(RotaryEmbedding):
    def __init__(self, kv_channels, rotary_percent=1.0, rotary_interleaved=False, seq_len_interpolation_factor=None,
                 rotary_base=10000.0, scaling_factor=1.0, original_max_position_embeddings=4096, beta_fast=32.0, beta_slow=1.0,
                 mscale=1.0, mscale_all_dim=0.0, use_cpu_initialization=False, cp_group=None):
        super().__init__(kv_channels, rotary_percent, rotary_interleaved, seq_len_interpolation_factor, rotary_base,
                         use_cpu_initialization=use_cpu_initialization, cp_group=cp_group)
        self.dim = kv_channels if rotary_percent >= 1.0 else int(kv_channels * rotary_percent)
        self.base, self.scaling_factor = rotary_base, scaling_factor
        self.orig_max, self.beta_fast, self.beta_slow = original_max_position_embeddings, beta_fast, beta_slow
        self.mscale, self.mscale_all_dim = mscale, mscale_all_dim
        dev = self.inv_freq.device
        extra = 1.0 / (rotary_base ** (torch.arange(0, self.dim, 2, dtype=torch.float32, device=dev) / self.dim))
        inter = extra / scaling_factor
        lo, hi = _find_correction_range(beta_fast, beta_slow, self.dim, rotary_base, original_max_position_embeddings)
        mask = 1.0 - _linear_ramp_mask(lo, hi, self.dim // 2, dev)
        self.inv_freq = inter * (1 - mask) + extra * mask

    @torch.no_grad()
    def forward(self, max_seq_len, offset=0, packed_seq=False, cp_group=None):
        emb = super().forward(max_seq_len, offset, packed_seq, cp_group)
        m = yarn_get_mscale(self.scaling_factor, self.mscale) / yarn_get_mscale(self.scaling_factor, self.mscale_all_dim)
        return emb, float(m)
