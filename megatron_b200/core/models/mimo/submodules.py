"""Per-modality encoder bundles (reference ``models/mimo/submodules/{base,vision,audio}.py``).

A ``ModalitySubmodules`` owns named encoders, optional decoders and the input/output projections into / out of the language
model's hidden size.  ``forward(encoder_inputs)`` → ``[n_embeddings, hidden]``: encode with every encoder (``encoder_inputs[name]``
is that encoder's kwargs), flatten each output to token rows, concatenate, project."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ...transformer.spec_utils import ModuleSpec, build_module


class ModalitySubmodules(torch.nn.Module):
    modality = "generic"

    def __init__(self, encoders: Optional[Dict[str, torch.nn.Module]] = None, decoders: Optional[Dict[str, torch.nn.Module]] = None,
                 input_projections: Optional[List[torch.nn.Module]] = None, output_projections: Optional[List[torch.nn.Module]] = None,
                 submodules: Optional[dict] = None, **kwargs):
        super().__init__()
        if submodules is not None:       # built from a ModuleSpec: {"encoders": {name: spec}, "input_projections": [spec], ...}
            encoders = {k: build_module(v) for k, v in (submodules.get("encoders") or {}).items()}
            decoders = {k: build_module(v) for k, v in (submodules.get("decoders") or {}).items()}
            input_projections = [build_module(s) for s in (submodules.get("input_projections") or [])]
            output_projections = [build_module(s) for s in (submodules.get("output_projections") or [])]
        self.encoders = torch.nn.ModuleDict(encoders or {})
        self.decoders = torch.nn.ModuleDict(decoders or {})
        self.input_projections = torch.nn.ModuleList(input_projections or [])
        self.output_projections = torch.nn.ModuleList(output_projections or [])

    @classmethod
    def from_spec(cls, spec: ModuleSpec, **kwargs):
        return build_module(spec, **kwargs)

    def encode(self, encoder_inputs: Dict[str, dict]) -> List[torch.Tensor]:
        outs = []
        for name, enc in self.encoders.items():
            if name not in encoder_inputs:
                continue
            y = enc(**encoder_inputs[name])
            outs.append(y.reshape(-1, y.shape[-1]))       # [b, n, h] → rows, sample-major
        return outs

    def combine_embeddings(self, embeddings: List[torch.Tensor]) -> torch.Tensor:
        if not embeddings:
            raise ValueError(f"no encoder of modality '{self.modality}' received inputs")
        return embeddings[0] if len(embeddings) == 1 else torch.cat(embeddings, dim=0)

    def project_embeddings(self, x: torch.Tensor, is_input: bool = True) -> torch.Tensor:
        for proj in (self.input_projections if is_input else self.output_projections):
            y = proj(x)
            x = y[0] if isinstance(y, tuple) else y
        return x

    def decode(self, embeddings: torch.Tensor, data_batch: Optional[dict] = None):
        out = {}
        x = self.project_embeddings(embeddings, is_input=False)
        for name, dec in self.decoders.items():
            out[name] = dec(x, **((data_batch or {}).get(name) or {}))
        return out

    def forward(self, encoder_inputs: Dict[str, dict]) -> Optional[torch.Tensor]:
        embs = self.encode(encoder_inputs)
        if not embs:
            return None
        return self.project_embeddings(self.combine_embeddings(embs), is_input=True)


class VisionModalitySubmodules(ModalitySubmodules):
    modality = "vision"


class AudioModalitySubmodules(ModalitySubmodules):
    modality = "audio"
