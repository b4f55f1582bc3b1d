"""RADIO vision tower (reference ``models/vision/radio.py``): a ViT with several class/summary tokens plus register tokens in front of the patch
tokens, a linear (not conv) patch embedder on pixel-unshuffled patches, and interpolatable absolute position embeddings so one checkpoint
serves several input resolutions.  Output: ``[b, n_summary + n_patches, h]`` (registers are dropped, as in the reference's ``class_token_len`` slicing)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec
from ...transformer.torch_norm import FusedNorm
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig


class RADIOViTModel(MegatronModule):
    def __init__(self, transformer_config: TransformerConfig, transformer_layer_spec: ModuleSpec, patch_dim: int = 16, img_h: int = 224, img_w: int = 224,
                 max_img_h: int = 2048, max_img_w: int = 2048, class_token_len: int = 8, num_registers: int = 0, add_class_token: bool = True,
                 ln_post: bool = False, pg_collection=None):
        super().__init__(config=transformer_config)
        c = transformer_config
        self.patch_dim, self.img_h, self.img_w = patch_dim, img_h, img_w
        self.class_token_len = class_token_len if add_class_token else 0
        self.num_registers = num_registers
        self.max_grid = (max_img_h // patch_dim, max_img_w // patch_dim)
        dev = "cpu" if (c.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.embedder = torch.nn.Linear(3 * patch_dim * patch_dim, c.hidden_size, bias=False, device=dev, dtype=c.params_dtype)
        self.position_embeddings = torch.nn.Parameter(0.02 * torch.randn(1, self.max_grid[0] * self.max_grid[1], c.hidden_size, device=dev, dtype=c.params_dtype))
        n_special = self.class_token_len + num_registers
        self.class_token = torch.nn.Parameter(torch.randn(n_special, c.hidden_size, device=dev, dtype=c.params_dtype)) if n_special else None
        self.decoder = TransformerBlock(config=c, spec=transformer_layer_spec, pre_process=True, post_process=False, pg_collection=pg_collection)
        self.ln_post = FusedNorm(c, c.hidden_size, eps=c.layernorm_epsilon) if ln_post else None

    def set_input_tensor(self, input_tensor):
        self.decoder.set_input_tensor(input_tensor)

    def _position_embeddings(self, gh: int, gw: int) -> torch.Tensor:
        """Crop (training-time augmentation is skipped) or bilinearly resize the [max_gh, max_gw] table to the current grid."""
        H, W = self.max_grid
        pe = self.position_embeddings.view(1, H, W, -1).permute(0, 3, 1, 2)
        if (gh, gw) != (H, W):
            pe = F.interpolate(pe.float(), size=(gh, gw), mode="bilinear", align_corners=False).to(self.position_embeddings.dtype)
        return pe.flatten(2).transpose(1, 2)           # [1, gh*gw, h]

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, _, ih, iw = x.shape
        p = self.patch_dim
        gh, gw = ih // p, iw // p
        patches = x.to(self.embedder.weight.dtype).unfold(2, p, p).unfold(3, p, p)            # [b, 3, gh, gw, p, p]
        patches = patches.permute(0, 2, 3, 1, 4, 5).reshape(b, gh * gw, 3 * p * p)
        tok = self.embedder(patches) + self._position_embeddings(gh, gw)
        if self.class_token is not None:
            tok = torch.cat([self.class_token.unsqueeze(0).expand(b, -1, -1), tok], dim=1)
        h = self.decoder(tok.permute(1, 0, 2).contiguous(), attention_mask).permute(1, 0, 2).contiguous()
        if self.ln_post is not None:
            h = self.ln_post(h)
        if self.num_registers:
            h = torch.cat([h[:, : self.class_token_len], h[:, self.class_token_len + self.num_registers :]], dim=1)
        return h
