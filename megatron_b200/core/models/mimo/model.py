"""Multi-in / multi-out model: modality encoders feeding one language model (reference ``models/mimo/model/base.py`` — ``MimoModel`` :26,
``align_embeddings_by_token_positions`` :146, ``forward`` :408).

``input_ids`` carries one placeholder token per modality embedding; the forward pass embeds the text, runs every hosted modality's
submodules, scatters their rows into the placeholder positions (sample-major order, vectorised — no host sync, so it is CUDA-graph safe)
and calls the language model with ``decoder_input``.  With ``module_to_grid_map`` a rank only builds the modules it hosts; encoder-only
ranks return the embeddings so the pipeline communicator (``MultiModulePipelineCommunicator``) can ship them to the language grid."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from ...transformer.module import MegatronModule
from ...transformer.spec_utils import build_module
from .config import MIMO_LANGUAGE_MODULE_KEY, MimoModelConfig


class MimoModel(MegatronModule):
    def __init__(self, mimo_config: MimoModelConfig, cp_group=None, tp_group=None):
        super().__init__(config=getattr(mimo_config.language_model_spec, "params", {}).get("config"))
        self.mimo_config = mimo_config
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.role = mimo_config.role_of(rank)
        self.special_token_ids = dict(mimo_config.special_token_ids)
        self.modality_submodules = torch.nn.ModuleDict(
            {name: build_module(spec) for name, spec in mimo_config.modality_submodules_spec.items() if self.role[name]})
        self.language_model = build_module(mimo_config.language_model_spec) if self.role[MIMO_LANGUAGE_MODULE_KEY] else None
        self.model_type = None

    # ---- helpers ----
    def set_input_tensor(self, input_tensor):
        if self.language_model is not None:
            self.language_model.set_input_tensor(input_tensor)

    def get_text_embeddings(self, input_ids: torch.Tensor, position_ids: Optional[torch.Tensor]) -> torch.Tensor:
        """[s, b, h] text embeddings with the placeholder ids mapped to token 0 (their rows are overwritten afterwards)."""
        ids = input_ids
        for tok in self.special_token_ids.values():
            ids = torch.where(ids == tok, torch.zeros_like(ids), ids)
        return self.language_model.embedding(input_ids=ids, position_ids=position_ids)

    @staticmethod
    def align_embeddings_by_token_positions(modality_embeddings: Dict[str, torch.Tensor], input_ids: torch.Tensor, special_token_ids: Dict[str, int],
                                            text_embeddings: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Scatter ``modality_embeddings[name]`` ([n, h], sample-major) into the positions of ``special_token_ids[name]`` → [s, b, h]."""
        if input_ids.dim() == 1:
            input_ids = input_ids[None]
        b, s = input_ids.shape
        ref = text_embeddings if text_embeddings is not None else next(iter(modality_embeddings.values()))
        h = ref.shape[-1]
        out = text_embeddings.transpose(0, 1).reshape(b * s, h) if text_embeddings is not None else ref.new_zeros(b * s, h)
        flat = input_ids.reshape(-1)
        for name, emb in modality_embeddings.items():
            if name == "text" or emb is None:
                continue
            mask = flat == special_token_ids[name]
            # k-th placeholder (row-major over [b, s]) takes the k-th embedding row; masked_scatter keeps everything on the device
            out = out.masked_scatter(mask[:, None], emb.to(out.dtype))
        return out.reshape(b, s, h).transpose(0, 1).contiguous()

    # ---- forward ----
    def forward(self, input_ids: torch.Tensor, position_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                loss_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None, modality_inputs: Optional[Dict[str, Dict[str, dict]]] = None,
                modality_embeddings: Optional[Dict[str, torch.Tensor]] = None, packing_kwargs: Optional[dict] = None):
        """``modality_inputs[modality][encoder_name]`` = kwargs of that encoder.  ``modality_embeddings`` are pre-computed rows received from
        encoder ranks (non-colocated placement).  Returns the language model output, or the embedding dict on encoder-only ranks."""
        embs: Dict[str, torch.Tensor] = dict(modality_embeddings or {})
        for name, sub in self.modality_submodules.items():
            if modality_inputs and name in modality_inputs and modality_inputs[name]:
                e = sub(modality_inputs[name])
                if e is not None:
                    embs[name] = e
        if self.language_model is None:
            return embs
        for name, e in embs.items():       # the count check is cheap insurance against silent misalignment (host sync only in debug mode)
            if __debug__ and not torch.cuda.is_available():
                n = int((input_ids == self.special_token_ids[name]).sum())
                if n != e.shape[0]:
                    raise ValueError(f"modality '{name}': {e.shape[0]} embeddings for {n} placeholder tokens")
        text = self.get_text_embeddings(input_ids, position_ids)
        sp = self.language_model.config.sequence_parallel
        if sp:       # the embedding already scattered along the sequence: merge on full sequences, then scatter again
            from ...tensor_parallel.mappings import gather_from_sequence_parallel_region, scatter_to_sequence_parallel_region

            text = gather_from_sequence_parallel_region(text, tensor_parallel_output_grad=False)
        combined = self.align_embeddings_by_token_positions(embs, input_ids, self.special_token_ids, text_embeddings=text)
        if sp:
            combined = scatter_to_sequence_parallel_region(combined)
        kw = {}
        if loss_mask is not None:
            kw["loss_mask"] = loss_mask
        return self.language_model(input_ids=None, position_ids=position_ids, attention_mask=attention_mask, decoder_input=combined, labels=labels, **kw)
