"""CLIP / SigLIP style ViT image encoder (reference ``models/vision/clip_vit_model.py``).

image [b, 3, H, W] → conv patchify → (+ class token) + learned positions → pre-LN → TransformerBlock → [b, n_tokens, h]."""
from __future__ import annotations

from typing import Optional

import torch

from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec
from ...transformer.torch_norm import FusedNorm
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_config import TransformerConfig


def get_num_image_embeddings(img_h: int, img_w: int, patch_dim: int, vision_model_type: str = "clip", disable_vision_class_token: bool = False,
                             class_token_len: int = 1, pixel_shuffle: bool = False, use_tile_tags: bool = False) -> int:
    n = (img_h // patch_dim) * (img_w // patch_dim)
    if vision_model_type == "clip" and not disable_vision_class_token:
        n += class_token_len
    if pixel_shuffle:
        n = n // 4
    return n + (5 if use_tile_tags else 0)


class CLIPViTModel(MegatronModule):
    def __init__(self, transformer_config: TransformerConfig, transformer_layer_spec: ModuleSpec, ln_pre_impl=FusedNorm, ln_post_impl=None,
                 add_class_token: bool = True, class_token_len: int = 1, patch_dim: int = 14, img_h: int = 336, img_w: int = 336,
                 model_subtype: str = "clip", pg_collection=None, vp_stage=None):
        super().__init__(config=transformer_config)
        c = transformer_config
        self.visual_hidden_size, self.patch_dim, self.img_h, self.img_w = c.hidden_size, patch_dim, img_h, img_w
        assert img_h % patch_dim == 0 and img_w % patch_dim == 0
        self.num_patches = (img_h // patch_dim) * (img_w // patch_dim)
        self.add_class_token, self.class_token_len = add_class_token, class_token_len
        self.seq_length = self.num_patches + (class_token_len if add_class_token else 0)
        dev = "cpu" if (c.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        conv_bias = model_subtype != "clip"
        self.conv1 = torch.nn.Conv2d(3, c.hidden_size, kernel_size=patch_dim, stride=patch_dim, bias=conv_bias, device=dev, dtype=c.params_dtype)
        self.position_embeddings = torch.nn.Embedding(self.seq_length, c.hidden_size, device=dev, dtype=c.params_dtype)
        self.register_buffer("position_ids", torch.arange(self.seq_length, device=dev).expand(1, -1), persistent=False)
        if add_class_token:
            self.class_token = torch.nn.Parameter(torch.randn(1, class_token_len, c.hidden_size, device=dev, dtype=c.params_dtype))
        self.ln_pre = ln_pre_impl(c, c.hidden_size, eps=c.layernorm_epsilon) if (ln_pre_impl is not None and model_subtype == "clip") else None
        self.decoder = TransformerBlock(config=c, spec=transformer_layer_spec, pre_process=True, post_process=False, pg_collection=pg_collection)
        # which norms exist follows the model family, as in the reference (clip_vit_model.py:87-113): CLIP normalises BEFORE the blocks, SigLIP after, InternViT neither
        self.ln_post = ln_post_impl(c, c.hidden_size, eps=c.layernorm_epsilon) if (ln_post_impl is not None and model_subtype == "siglip") else None
        self.model_type = None

    def set_input_tensor(self, input_tensor):
        self.decoder.set_input_tensor(input_tensor)

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self.conv1(x.to(self.conv1.weight.dtype))                 # [b, h, gh, gw]
        x = x.flatten(2).transpose(1, 2)                              # [b, patches, h]
        if self.add_class_token:
            x = torch.cat([self.class_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.position_embeddings(self.position_ids)
        if self.ln_pre is not None:
            x = self.ln_pre(x)
        x = x.permute(1, 0, 2).contiguous()                           # [s, b, h]
        x = self.decoder(x, attention_mask)
        x = x.permute(1, 0, 2).contiguous()
        if self.ln_post is not None:
            x = self.ln_post(x)
        return x
