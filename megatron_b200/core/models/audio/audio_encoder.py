"""Whisper-style audio encoder (reference ``models/audio/`` + ``models/mimo/submodules/audio.py``).

log-mel features [b, n_mels, frames] → conv1d(k3) → GELU → conv1d(k3, stride 2) → GELU → + sinusoidal positions → TransformerBlock
(bidirectional) → LayerNorm → [b, frames/2, h].  The transformer stack is the same ``TransformerBlock`` the language models use, so
it inherits tensor parallelism and the fused kernels."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec
from ...transformer.torch_norm import FusedNorm
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig


def get_num_audio_embeddings(num_frames: int, stride: int = 2) -> int:
    return (num_frames + stride - 1) // stride


def _sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


class AudioEncoderModel(MegatronModule):
    def __init__(self, transformer_config: TransformerConfig, transformer_layer_spec: ModuleSpec, n_mels: int = 80, max_frames: int = 3000,
                 pg_collection=None):
        super().__init__(config=transformer_config)
        c = transformer_config
        dev = "cpu" if (c.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.conv1 = torch.nn.Conv1d(n_mels, c.hidden_size, kernel_size=3, padding=1, device=dev, dtype=c.params_dtype)
        self.conv2 = torch.nn.Conv1d(c.hidden_size, c.hidden_size, kernel_size=3, stride=2, padding=1, device=dev, dtype=c.params_dtype)
        self.register_buffer("positions", _sinusoids(get_num_audio_embeddings(max_frames), c.hidden_size).to(dev), persistent=False)
        self.decoder = TransformerBlock(config=c, spec=transformer_layer_spec, pre_process=True, post_process=False, pg_collection=pg_collection)
        self.ln_post = FusedNorm(c, c.hidden_size, eps=c.layernorm_epsilon)

    def set_input_tensor(self, input_tensor):
        self.decoder.set_input_tensor(input_tensor)

    def forward(self, mel: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = F.gelu(self.conv1(mel.to(self.conv1.weight.dtype)))
        x = F.gelu(self.conv2(x))                       # [b, h, frames/2]
        x = x.transpose(1, 2)
        x = x + self.positions[: x.shape[1]].to(x.dtype)
        x = x.permute(1, 0, 2).contiguous()             # [s, b, h]
        x = self.decoder(x, attention_mask)
        return self.ln_post(x.permute(1, 0, 2).contiguous())
