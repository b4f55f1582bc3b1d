"""LLaVA-style vision-language model (reference ``models/multimodal/llava_model.py``).

images → CLIPViTModel → MultimodalProjector → spliced into the text embedding sequence at ``image_token_index`` placeholders →
GPTModel.  Labels / loss mask are expanded consistently (image positions are never predicted)."""
from __future__ import annotations

from typing import Optional

import torch

from ...transformer.mlp import MLPSubmodules
from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_config import TransformerConfig
from ..gpt.gpt_model import GPTModel
from ..vision.clip_vit_model import CLIPViTModel, get_num_image_embeddings
from ..vision.multimodal_projector import MultimodalProjector

IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN_INDEX = -200
IMAGE_TOKEN = "<image>"


class LLaVAModel(MegatronModule):
    def __init__(self, language_transformer_config: TransformerConfig, language_transformer_layer_spec: ModuleSpec, language_vocab_size: int,
                 language_max_sequence_length: int, vision_transformer_config: TransformerConfig, vision_transformer_layer_spec: ModuleSpec,
                 drop_vision_class_token: bool, vision_projection_config: TransformerConfig, vision_projection_layer_spec: MLPSubmodules,
                 vision_projection_type: str = "mlp", allow_missing_vision_projection_checkpoint: bool = False, parallel_output: bool = True,
                 share_embeddings_and_output_weights: bool = False, language_position_embedding_type: str = "learned_absolute",
                 language_rotary_percent: float = 1.0, pre_process: bool = True, post_process: bool = True, add_encoder: bool = True,
                 add_decoder: bool = True, img_h: int = 336, img_w: int = 336, patch_dim: int = 14, language_rotary_base: int = 10000,
                 image_token_index: int = DEFAULT_IMAGE_TOKEN_INDEX, pg_collection=None, vp_stage=None):
        super().__init__(config=language_transformer_config)
        self.pre_process, self.post_process, self.add_encoder, self.add_decoder = pre_process, post_process, add_encoder, add_decoder
        self.image_token_index = image_token_index
        self.language_model = None
        if add_decoder:
            self.language_model = GPTModel(
                config=language_transformer_config, transformer_layer_spec=language_transformer_layer_spec, vocab_size=language_vocab_size,
                max_sequence_length=language_max_sequence_length, parallel_output=parallel_output,
                share_embeddings_and_output_weights=share_embeddings_and_output_weights, position_embedding_type=language_position_embedding_type,
                rotary_percent=language_rotary_percent, pre_process=pre_process, post_process=post_process, rotary_base=language_rotary_base,
                scatter_embedding_sequence_parallel=False, pg_collection=pg_collection, vp_stage=vp_stage,
            )
            self.share_embeddings_and_output_weights = self.language_model.share_embeddings_and_output_weights
        self.vision_model = self.vision_projection = None
        self._drop_vision_class_token = drop_vision_class_token
        if add_encoder:
            self.vision_model = CLIPViTModel(vision_transformer_config, vision_transformer_layer_spec, img_h=img_h, img_w=img_w, patch_dim=patch_dim)
            self.vision_projection = MultimodalProjector(vision_projection_config, vision_projection_layer_spec, vision_projection_type,
                                                         vision_transformer_config.hidden_size)
        self.img_seq_len = get_num_image_embeddings(img_h, img_w, patch_dim, "clip", drop_vision_class_token, 1)
        self.model_type = None

    def shared_embedding_or_output_weight(self):
        return self.language_model.shared_embedding_or_output_weight() if self.add_decoder else None

    def set_input_tensor(self, input_tensor):
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        if self.add_encoder and self.add_decoder:
            self.vision_model.set_input_tensor(input_tensor[0])
        elif self.add_encoder:
            self.vision_model.set_input_tensor(input_tensor[0])
        elif self.pre_process:
            self.encoder_hidden_state = input_tensor[0]
        else:
            self.language_model.set_input_tensor(input_tensor[0])

    def freeze(self, freeze_language_model: bool, freeze_vision_model: bool, freeze_vision_projection: bool):
        mods = []
        if freeze_language_model and self.language_model is not None:
            mods.append(self.language_model)
        if freeze_vision_model and self.vision_model is not None:
            mods.append(self.vision_model)
        if freeze_vision_projection and self.vision_projection is not None:
            mods.append(self.vision_projection)
        for m in mods:
            for p in m.parameters():
                p.requires_grad = False

    def _preprocess_data(self, image_embeddings, language_embeddings, input_ids, loss_mask, labels, num_image_tiles):
        """Splice image embeddings [img_seq, n_tiles, h] into text embeddings [b, s, h] at the placeholder positions.
        Returns (combined [s', b, h], labels [b, s'], loss_mask [b, s'])."""
        b, s = input_ids.shape
        h = language_embeddings.shape[-1]
        img_seq = image_embeddings.shape[0]
        is_img = input_ids == self.image_token_index
        n_img_per_sample = is_img.sum(dim=1)
        if num_image_tiles is None:
            num_image_tiles = torch.ones(int(n_img_per_sample.sum()), dtype=torch.long, device=input_ids.device)
        tiles = num_image_tiles.tolist()
        out_emb, out_lab, out_msk, lens = [], [], [], []
        tile_cursor = img_cursor = 0
        for i in range(b):
            pieces_e, pieces_l, pieces_m = [], [], []
            start = 0
            positions = is_img[i].nonzero(as_tuple=True)[0].tolist()
            for pos in positions + [s]:
                if pos > start:
                    pieces_e.append(language_embeddings[i, start:pos])
                    if labels is not None:
                        pieces_l.append(labels[i, start:pos])
                        pieces_m.append(loss_mask[i, start:pos])
                if pos < s:
                    nt = tiles[img_cursor]
                    img = image_embeddings[:, tile_cursor : tile_cursor + nt].permute(1, 0, 2).reshape(nt * img_seq, h)
                    pieces_e.append(img.to(language_embeddings.dtype))
                    if labels is not None:
                        pieces_l.append(torch.full((nt * img_seq,), IGNORE_INDEX, dtype=labels.dtype, device=labels.device))
                        pieces_m.append(torch.zeros(nt * img_seq, dtype=loss_mask.dtype, device=loss_mask.device))
                    tile_cursor += nt
                    img_cursor += 1
                start = pos + 1
            e = torch.cat(pieces_e, dim=0)
            out_emb.append(e)
            lens.append(e.shape[0])
            if labels is not None:
                out_lab.append(torch.cat(pieces_l))
                out_msk.append(torch.cat(pieces_m))
        L = max(lens)
        emb = language_embeddings.new_zeros(b, L, h)
        lab = msk = None
        if labels is not None:
            lab = torch.full((b, L), IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
            msk = torch.zeros(b, L, dtype=loss_mask.dtype, device=loss_mask.device)
        for i in range(b):
            emb[i, : lens[i]] = out_emb[i]
            if labels is not None:
                lab[i, : lens[i]] = out_lab[i]
                msk[i, : lens[i]] = out_msk[i]
        return emb.transpose(0, 1).contiguous(), lab, msk

    def forward(self, images: torch.Tensor, input_ids: torch.Tensor, position_ids: torch.Tensor, attention_mask: torch.Tensor = None,
                labels: Optional[torch.Tensor] = None, loss_mask: Optional[torch.Tensor] = None, inference_context=None,
                num_image_tiles: Optional[torch.Tensor] = None, image_token_index: Optional[int] = None, runtime_gather_output=None, *, inference_params=None):
        """images [n_tiles, 3, H, W].  Returns (per-token loss [b, s'], loss_mask [b, s']) when labels are given, else logits."""
        inference_context = inference_context or inference_params
        if image_token_index is not None:
            self.image_token_index = image_token_index
        use_kv = inference_context is not None and "image_tokens_count" in inference_context.key_value_memory_dict
        has_images = images is not None and images.shape[0] > 0
        image_embeddings = None
        if use_kv or not has_images:
            pass
        elif self.add_encoder:
            ie = self.vision_model(images)                              # [n, img_seq(+cls), hv]
            if self._drop_vision_class_token:
                ie = ie[:, self.vision_model.class_token_len :, :]
            ie = ie.permute(1, 0, 2).contiguous()                       # [img_seq, n, hv]
            image_embeddings = self.vision_projection(ie)               # [img_seq, n, h]
            if inference_context is not None:
                inference_context.key_value_memory_dict["image_tokens_count"] = image_embeddings.shape[0] * image_embeddings.shape[1]
        else:
            image_embeddings = self.encoder_hidden_state
        if not self.add_decoder:
            return image_embeddings, loss_mask
        combined = None
        new_labels, new_mask = labels, loss_mask
        if self.pre_process:
            ids_text = input_ids.clone()
            ids_text[ids_text == self.image_token_index] = 0
            le = self.language_model.embedding(input_ids=ids_text, position_ids=position_ids).transpose(0, 1).contiguous()  # [b, s, h]
            if image_embeddings is not None:
                if labels is not None and loss_mask is None:
                    loss_mask = torch.ones_like(labels, dtype=torch.float32)
                combined, new_labels, new_mask = self._preprocess_data(image_embeddings, le, input_ids, loss_mask, labels, num_image_tiles)
            else:
                combined = le.transpose(0, 1).contiguous()
            if self.config.sequence_parallel:
                from ... import tensor_parallel

                combined = tensor_parallel.scatter_to_sequence_parallel_region(combined)
        lab_for_lm = None
        if new_labels is not None:
            lab_for_lm = new_labels.clone()
            lab_for_lm[lab_for_lm == IGNORE_INDEX] = 0
        out = self.language_model(input_ids=None, position_ids=None, attention_mask=attention_mask, decoder_input=combined, labels=lab_for_lm,
                                  inference_context=inference_context, runtime_gather_output=runtime_gather_output)
        if labels is None or not self.post_process:
            return out
        return out, new_mask
