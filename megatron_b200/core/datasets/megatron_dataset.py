"""Dataset base classes (reference ``datasets/megatron_dataset.py``, ``blended_megatron_dataset_config.py``)."""
from __future__ import annotations

import hashlib
import json
from abc import ABC, abstractmethod
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy
import torch

from .utils import Split, normalize


@dataclass
class BlendedMegatronDatasetConfig:
    random_seed: int
    sequence_length: int
    blend: Optional[Tuple[List[str], Optional[List[float]]]] = None
    blend_per_split: Optional[List[Optional[Tuple[List[str], Optional[List[float]]]]]] = None
    multiple_validation_sets: Optional[bool] = None
    full_validation: Optional[bool] = None
    split: Optional[str] = None
    split_matrix: Optional[List[Tuple[float, float]]] = field(init=False, default=None)
    num_dataset_builder_threads: int = 1
    path_to_cache: Optional[str] = None
    mmap_bin_files: bool = True
    mock: bool = field(init=False, default=False)
    tokenizer: Optional[Any] = None
    mid_level_dataset_surplus: float = 0.005

    def __post_init__(self):
        if self.blend_per_split is not None and any(self.blend_per_split):
            assert self.blend is None, "blend and blend_per_split are incompatible"
            assert self.split is None, "split and blend_per_split are incompatible"
            assert len(self.blend_per_split) == len(Split)
        else:
            if self.blend is None:
                self.mock = True
                self.split = self.split or "1,1,1"
            assert self.split is not None, "split must be provided in absence of blend_per_split"
            self.split_matrix = convert_split_vector_to_split_matrix(parse_and_normalize_split(self.split))


def parse_and_normalize_split(split: str) -> List[float]:
    parts = [float(s) for s in split.replace("/", ",").split(",") if s.strip()]
    parts = parts + [0.0] * (len(Split) - len(parts))
    assert len(parts) == len(Split) and all(p >= 0 for p in parts)
    return normalize(parts)


def convert_split_vector_to_split_matrix(vector_a: List[float], vector_b: Optional[List[float]] = None) -> List[Optional[Tuple[float, float]]]:
    """[0.9,0.1,0] → [(0,0.9),(0.9,1.0),None]; with ``vector_b`` the per-split intersection."""
    if vector_b is None:
        vector_b = vector_a
    acc_a = [0.0] + list(numpy.cumsum(vector_a))
    acc_b = [0.0] + list(numpy.cumsum(vector_b))
    out = []
    for i in range(len(vector_a)):
        lo, hi = max(acc_a[i], acc_b[i]), min(acc_a[i + 1], acc_b[i + 1])
        out.append((float(lo), float(hi)) if hi > lo else None)
    return out


class LowLevelDataset:
    pass


class MegatronDataset(ABC, torch.utils.data.Dataset):
    """A (low-level dataset, index subset, split) triple with a content hash used for index caching."""

    def __init__(self, dataset, dataset_path: Optional[str], indices: numpy.ndarray, num_samples: Optional[int], index_split: Split, config):
        self.dataset, self.dataset_path, self.indices, self.num_samples = dataset, dataset_path, indices, num_samples
        self.index_split, self.config = index_split, config
        self.unique_identifiers = OrderedDict(
            **{"class": type(self).__name__, "dataset_path": dataset_path, "num_samples": num_samples, "index_split": index_split.name},
            **{k: getattr(config, k) for k in self._key_config_attributes() if hasattr(config, k)},
        )
        self.unique_description = json.dumps(self.unique_identifiers, indent=4, default=lambda o: getattr(o, "unique_identifiers", str(o)))
        self.unique_description_hash = hashlib.md5(self.unique_description.encode("utf-8"), usedforsecurity=False).hexdigest()
        self.built_anew_on_cache_miss = False

    @staticmethod
    def numel_low_level_dataset(low_level_dataset) -> int:
        raise NotImplementedError

    @staticmethod
    def build_low_level_dataset(dataset_path: str, config):
        raise NotImplementedError

    @staticmethod
    def _key_config_attributes() -> List[str]:
        return ["random_seed", "sequence_length", "split", "split_matrix", "tokenizer"]

    @abstractmethod
    def __len__(self) -> int:
        ...

    @abstractmethod
    def __getitem__(self, idx: int) -> Dict[str, Union[torch.Tensor, numpy.ndarray]]:
        ...
