"""Mock image+text dataset for VLM pre-training smoke runs (reference ``datasets/multimodal_dataset.py:12-62``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from .gpt_dataset import GPTDatasetConfig, MockGPTDataset


@dataclass
class MultimodalDatasetConfig(GPTDatasetConfig):
    image_h: Optional[int] = None
    image_w: Optional[int] = None
    preprocess_func: Optional[Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]]] = None   # e.g. insert image-token placeholders

    def __post_init__(self) -> None:
        super().__post_init__()
        assert self.image_h is not None and self.image_w is not None, "image_h and image_w are required"


class MockMultimodalDataset(MockGPTDataset):
    """The mock GPT sample plus a deterministic pseudo-random image ``[3, H, W]`` per index."""

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        sample = super().__getitem__(idx)
        g = torch.Generator().manual_seed(int(self.config.random_seed) * 1_000_003 + int(idx if idx is not None else 0))
        sample["image"] = torch.rand(3, self.config.image_h, self.config.image_w, generator=g)
        if self.config.preprocess_func is not None:
            sample = self.config.preprocess_func(sample)
        return sample
