"""Builds train/valid/test datasets from a blend spec (reference ``datasets/blended_megatron_dataset_builder.py:30``)."""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Type, Union

import numpy
import torch

from .blended_dataset import BlendedDataset
from .megatron_dataset import BlendedMegatronDatasetConfig, MegatronDataset
from .utils import Split, normalize


class BlendedMegatronDatasetBuilder:
    def __init__(self, cls: Type[MegatronDataset], sizes: List[Optional[int]], is_built_on_rank: Callable, config: BlendedMegatronDatasetConfig):
        self.cls, self.sizes, self.is_built_on_rank, self.config = cls, sizes, is_built_on_rank, config
        if torch.distributed.is_initialized() and torch.distributed.get_rank() == 0:
            assert is_built_on_rank(), "is_built_on_rank must return True when global rank = 0"

    def build(self) -> List[Optional[Union[MegatronDataset, BlendedDataset]]]:
        cfg = self.config
        if cfg.mock:
            return self._build_splits(None, self.sizes)
        if cfg.blend is not None:
            prefixes, weights = cfg.blend
            if len(prefixes) == 1 and weights is None:
                return self._build_splits(prefixes[0], self.sizes)
            return self._build_blend(prefixes, weights, self.sizes, [True] * len(Split))
        out: List = [None] * len(Split)
        for i in range(len(Split)):
            spec = cfg.blend_per_split[i]
            if spec is None:
                continue
            only = [self.sizes[j] if j == i else None for j in range(len(Split))]
            mask = [j == i for j in range(len(Split))]
            prefixes, weights = spec
            res = self._build_splits(prefixes[0], only, force_full=mask) if (len(prefixes) == 1 and weights is None) else self._build_blend(prefixes, weights, only, mask, force_full=True)
            out[i] = res[i]
        return out

    def _build_blend(self, prefixes, weights, sizes, mask, force_full=False):
        surplus = 1.0 + self.config.mid_level_dataset_surplus
        if weights is None:
            per = [[None] * len(Split) for _ in prefixes]
        else:
            w = normalize(weights)
            per = [[None if s is None else int(math.ceil(math.ceil(s * wi) * surplus)) for s in sizes] for wi in w]
        mids = [self._build_splits(p, ps, force_full=mask if force_full else None) for p, ps in zip(prefixes, per)]
        out = []
        for i in range(len(Split)):
            parts = [m[i] for m in mids]
            if any(p is None for p in parts) or not mask[i]:
                out.append(None)
                continue
            if weights is None:
                # blend by the components' own sizes; a requested size caps it (reference blended_megatron_dataset_builder.py:205-216)
                lens = [len(p) for p in parts]
                size_i = min(sizes[i], sum(lens)) if sizes[i] is not None else None
                out.append(self._guard(BlendedDataset, parts, lens, size_i, self.config))
            else:
                # the blend serves Σ_d ceil(target · w_d) samples — the per-dataset targets are rounded UP one by one, so the blend can be a few samples longer
                # than the request (reference :196-200); matching it keeps the blending indices (and therefore the data order) identical
                w = normalize(weights)
                size_i = None if sizes[i] is None else sum(int(math.ceil(sizes[i] * wi)) for wi in w)
                out.append(self._guard(BlendedDataset, parts, w, size_i, self.config))
        return out

    def _build_splits(self, dataset_path: Optional[str], sizes: List[Optional[int]], force_full=None):
        cfg = self.config
        low = self.cls.build_low_level_dataset(dataset_path, cfg) if self.is_built_on_rank() or True else None
        n = self.cls.numel_low_level_dataset(low)
        out = []
        for i, split in enumerate(Split):
            if force_full is not None:
                bounds = (0.0, 1.0) if force_full[i] else None
            else:
                bounds = cfg.split_matrix[i] if cfg.split_matrix is not None else None
            if bounds is None or (sizes[i] is None and cfg.mock is False and force_full is None and False):
                out.append(None)
                continue
            lo, hi = int(round(bounds[0] * n)), int(round(bounds[1] * n))
            idx = numpy.arange(lo, hi, dtype=numpy.int32)
            if len(idx) == 0:
                out.append(None)
                continue
            out.append(self._guard(self.cls, low, dataset_path, idx, sizes[i], split, cfg))
        return out

    def _guard(self, cls, *args):
        """Build on rank 0 first (populates the index cache), then on the other ranks."""
        if torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
            ds = None
            if rank == 0 and self.is_built_on_rank():
                ds = cls(*args)
            if getattr(self.config, "path_to_cache", None):
                torch.distributed.barrier()
            if rank != 0 and self.is_built_on_rank():
                ds = cls(*args)
            return ds
        return cls(*args)
