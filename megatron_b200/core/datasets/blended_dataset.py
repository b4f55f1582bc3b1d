"""Weighted interleave of several datasets (reference ``datasets/blended_dataset.py``)."""
from __future__ import annotations

import hashlib
import json
from collections import OrderedDict
from typing import Dict, List, Optional, Union

import numpy
import torch

from . import helpers
from .megatron_dataset import MegatronDataset
from .utils import normalize


class BlendedDataset(torch.utils.data.Dataset):
    def __init__(self, datasets: List[MegatronDataset], weights: List[Union[int, float]], size: Optional[int], config):
        assert len(datasets) == len(weights) and len(datasets) < 32767
        assert all(type(d) == type(datasets[0]) for d in datasets)
        if size is None and any(isinstance(w, float) for w in weights):
            raise AssertionError("weights must be sizes (ints) when the blend size is not given")
        self.datasets, self.split, self.weights, self.size, self.config = datasets, datasets[0].index_split, weights, size, config
        ident = OrderedDict(**{"class": type(self).__name__, "datasets": [d.unique_identifiers for d in datasets], "split": self.split.name,
                               "weights": list(weights), "size": size})
        self.unique_description = json.dumps(ident, indent=4, default=str)
        self.unique_description_hash = hashlib.md5(self.unique_description.encode("utf-8"), usedforsecurity=False).hexdigest()
        self.dataset_index, self.dataset_sample_index = self._build_indices()

    def __len__(self) -> int:
        return self.dataset_index.shape[0]

    def __getitem__(self, idx: int) -> Dict[str, Union[int, numpy.ndarray]]:
        d, s = int(self.dataset_index[idx]), int(self.dataset_sample_index[idx])
        return {"dataset_id": d, **self.datasets[d][s]}

    def _build_indices(self):
        n = len(self.datasets)
        if self.size is not None:
            di = numpy.zeros(self.size, dtype=numpy.int16)
            dsi = numpy.zeros(self.size, dtype=numpy.int64)
            helpers.build_blending_indices(di, dsi, normalize(self.weights), n, self.size, False)
        else:
            total = int(sum(self.weights))
            di = numpy.zeros(total, dtype=numpy.int16)
            dsi = numpy.zeros(total, dtype=numpy.int64)
            helpers.build_exhaustive_blending_indices(di, dsi, [int(w) for w in self.weights], n)
        for i, d in enumerate(self.datasets):
            need = int(dsi[di == i].max()) + 1 if (di == i).any() else 0
            if need > len(d):
                raise IndexError(f"blend needs {need} samples from dataset {i} which only has {len(d)}")
        return di, dsi
