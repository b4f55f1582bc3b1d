"""GPT pre-training dataset (reference ``datasets/gpt_dataset.py:141-998``).

Three index arrays turn a corpus of documents into fixed-length samples:
``document_index`` (epochs × shuffled doc ids) → ``sample_index`` (native ``build_sample_idx``:
where each ``seq_length+1`` window starts) → ``shuffle_index`` (sample order).  They are cached
as ``.npy`` under ``path_to_cache`` keyed by the dataset's content hash.
"""
from __future__ import annotations

import logging
import os
import time
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy
import torch

from . import helpers
from .indexed_dataset import IndexedDataset
from .megatron_dataset import BlendedMegatronDatasetConfig, MegatronDataset
from .utils import Split

logger = logging.getLogger(__name__)
_PAD_TOKEN_ID = -1


@dataclass
class GPTDatasetConfig(BlendedMegatronDatasetConfig):
    reset_position_ids: Optional[bool] = None
    reset_attention_mask: Optional[bool] = None
    eod_mask_loss: Optional[bool] = None
    create_attention_mask: bool = True
    drop_last_partial_validation_sequence: bool = True
    add_extra_token_to_sequence: bool = True
    object_storage_cache_path: Optional[str] = None

    def __post_init__(self):
        super().__post_init__()
        assert self.tokenizer is not None
        assert self.reset_position_ids is not None and self.reset_attention_mask is not None and self.eod_mask_loss is not None


class GPTDataset(MegatronDataset):
    def __init__(self, indexed_dataset: IndexedDataset, dataset_path: Optional[str], indexed_indices: numpy.ndarray, num_samples: Optional[int],
                 index_split: Split, config: GPTDatasetConfig):
        super().__init__(indexed_dataset, dataset_path, indexed_indices, num_samples, index_split, config)
        self.masks_and_position_ids_are_cacheable = not any([config.reset_position_ids, config.reset_attention_mask, config.eod_mask_loss])
        self._cached = None
        try:
            self._pad_token_id = self.config.tokenizer.pad
        except Exception:
            self._pad_token_id = _PAD_TOKEN_ID
        self.document_index, self.sample_index, self.shuffle_index = self._build_document_sample_shuffle_indices()

    @staticmethod
    def numel_low_level_dataset(low_level_dataset: IndexedDataset) -> int:
        return low_level_dataset.sequence_lengths.shape[0]

    @staticmethod
    def build_low_level_dataset(dataset_path: str, config: GPTDatasetConfig) -> IndexedDataset:
        return IndexedDataset(dataset_path, multimodal=False, mmap=config.mmap_bin_files)

    def __len__(self) -> int:
        return self.sample_index.shape[0] - 1

    def __getitem__(self, idx: Optional[int]) -> Dict[str, torch.Tensor]:
        if idx is None:
            text, _ = self._query(0)
        else:
            text, _ = self._query(idx)
        text = torch.from_numpy(text).long()
        if self.config.add_extra_token_to_sequence:
            tokens, labels = text[:-1].contiguous(), text[1:].contiguous()
        else:
            tokens = text
            labels = torch.roll(text, shifts=-1, dims=0)
            labels[-1] = self._pad_token_id
        if not self.masks_and_position_ids_are_cacheable or self._cached is None:
            am, lm, pid = _get_ltor_masks_and_position_ids(tokens, self.config.tokenizer.eod, self.config.reset_position_ids,
                                                           self.config.reset_attention_mask, self.config.eod_mask_loss, self.config.create_attention_mask)
            if self.masks_and_position_ids_are_cacheable:
                self._cached = (am, lm, pid)
        else:
            am, lm, pid = self._cached
        lm = lm.clone()
        lm[labels == self._pad_token_id] = 0.0
        tokens = tokens.clone()
        tokens[tokens == self._pad_token_id] = 0
        labels = labels.clone()
        labels[labels == self._pad_token_id] = 0
        if idx is None:
            lm = torch.zeros_like(lm)
        out = {"tokens": tokens, "labels": labels, "loss_mask": lm, "position_ids": pid}
        if self.config.create_attention_mask:
            out["attention_mask"] = am
        return out

    def _query(self, idx: int) -> Tuple[numpy.ndarray, numpy.ndarray]:
        idx = self.shuffle_index[idx]
        d0, o0 = self.sample_index[idx]
        d1, o1 = self.sample_index[idx + 1]
        extra = 1 if self.config.add_extra_token_to_sequence else 0
        parts, docs = [], []
        if d0 == d1:
            docs.append(self.document_index[d0])
            parts.append(self.dataset.get(self.document_index[d0], offset=int(o0), length=int(o1 - o0 + extra)))
        else:
            for i in range(d0, d1 + 1):
                docs.append(self.document_index[i])
                off = int(o0) if i == d0 else 0
                length = None if i < d1 else int(o1 + extra)
                parts.append(self.dataset.get(self.document_index[i], offset=off, length=length))
        text = numpy.concatenate(parts, dtype=numpy.int64)
        need = self.config.sequence_length + extra
        if len(text) < need:
            text = numpy.pad(text, (0, need - len(text)), constant_values=self._pad_token_id)
        return text, numpy.array(docs, dtype=numpy.int64)

    def _build_document_sample_shuffle_indices(self):
        cfg = self.config
        cache = cfg.path_to_cache
        if cache is None and not cfg.mock and self.dataset_path is not None:
            cache = os.path.join(os.path.dirname(self.dataset_path), "cache", f"{type(self).__name__}_indices")
        names = {}
        if cache:
            base = os.path.join(cache, f"{self.unique_description_hash}-{type(self).__name__}-{self.index_split.name}")
            names = {k: f"{base}-{k}.npy" for k in ("document_index", "sample_index", "shuffle_index")}
            if all(os.path.isfile(p) for p in names.values()):
                return tuple(numpy.load(names[k], allow_pickle=True, mmap_mode="r") for k in ("document_index", "sample_index", "shuffle_index"))
        t0 = time.time()
        rng = numpy.random.RandomState(cfg.random_seed)
        sizes = self.dataset.sequence_lengths
        tokens_per_epoch = int(numpy.sum(sizes[self.indices]))
        extra = 1 if cfg.add_extra_token_to_sequence else 0
        seq = cfg.sequence_length
        if self.num_samples is None:
            num_epochs = 1
        else:
            num_epochs, tokens = 1, tokens_per_epoch
            while (tokens - extra) // seq < self.num_samples:
                num_epochs += 1
                tokens += tokens_per_epoch
        # the last epoch is shuffled separately if it contributes < 80% of an epoch's samples
        separate_final = False
        if num_epochs > 1 and self.num_samples is not None:
            before = ((num_epochs - 1) * tokens_per_epoch - extra) // seq
            from_final = self.num_samples - before
            per_epoch = (tokens_per_epoch - extra) // seq
            separate_final = from_final < int(0.80 * per_epoch)
        doc_idx = _build_document_index(self.indices, num_epochs, rng, separate_final)
        drop_last = True if self.index_split != Split.valid else cfg.drop_last_partial_validation_sequence
        sample_index = helpers.build_sample_idx(sizes, doc_idx, seq, num_epochs, tokens_per_epoch, drop_last, cfg.add_extra_token_to_sequence)
        n = sample_index.shape[0] - 1
        if separate_final:
            n_first = ((num_epochs - 1) * tokens_per_epoch - extra) // seq
            shuffle = _build_shuffle_index(n_first, n, rng)
        else:
            shuffle = _build_shuffle_index(n, n, rng)
        if names and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0 or True):
            os.makedirs(cache, exist_ok=True)
            numpy.save(names["document_index"], doc_idx, allow_pickle=True)
            numpy.save(names["sample_index"], sample_index, allow_pickle=True)
            numpy.save(names["shuffle_index"], shuffle, allow_pickle=True)
        self.built_anew_on_cache_miss = True
        logger.info("built GPT indices for %s split in %.2fs (%d samples, %d epochs)", self.index_split.name, time.time() - t0, n, num_epochs)
        return doc_idx, sample_index, shuffle


def _build_document_index(documents: numpy.ndarray, num_epochs: int, rng: numpy.random.RandomState, separate_final_epoch: bool) -> numpy.ndarray:
    if not separate_final_epoch or num_epochs == 1:
        idx = numpy.tile(numpy.asarray(documents, dtype=numpy.int32), num_epochs)
        rng.shuffle(idx)
        return idx
    first = _build_document_index(documents, num_epochs - 1, rng, False)
    last = _build_document_index(documents, 1, rng, False)
    return numpy.concatenate((first, last))


def _build_shuffle_index(num_samples: int, total_size: int, rng: numpy.random.RandomState) -> numpy.ndarray:
    dt = numpy.uint32 if total_size < numpy.iinfo(numpy.uint32).max - 1 else numpy.int64
    a = numpy.arange(0, num_samples, dtype=dt)
    rng.shuffle(a)
    if num_samples == total_size:
        return a
    b = numpy.arange(num_samples, total_size, dtype=dt)
    rng.shuffle(b)
    return numpy.concatenate((a, b))


def _get_ltor_masks_and_position_ids(data: torch.Tensor, eod_token: int, reset_position_ids: bool, reset_attention_mask: bool, eod_mask_loss: bool,
                                     create_attention_mask: bool):
    """Left-to-right masks; optionally restart positions / block attention at document boundaries."""
    n = data.numel()
    am = torch.tril(torch.ones((n, n))).unsqueeze(0) if create_attention_mask else None
    lm = torch.ones(n, dtype=torch.float)
    if eod_mask_loss:
        lm[data == eod_token] = 0.0
    pid = torch.arange(n, dtype=torch.long)
    if reset_position_ids or reset_attention_mask:
        pid = pid.clone()
        prev = 0
        for i in (data == eod_token).nonzero().flatten().tolist():
            if reset_attention_mask and am is not None:
                am[0, (i + 1):, : (i + 1)] = 0
            if reset_position_ids:
                pid[(i + 1):] -= i + 1 - prev
                prev = i + 1
    if am is not None:
        am = am < 0.5
    return am, lm, pid


class MockGPTLowLevelDataset:
    """Deterministic synthetic corpus: 100k documents of random length and random tokens."""

    seed: int = 0
    size: int = 100000
    max_sequence_length: int = 4096

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer
        rng = numpy.random.default_rng(seed=self.seed)
        self.sequence_lengths = rng.integers(low=1, high=self.max_sequence_length, size=self.size, dtype=numpy.int32)

    def __len__(self):
        return self.size

    def __getitem__(self, idx: int) -> numpy.ndarray:
        length = int(self.sequence_lengths[idx])
        sample = numpy.int64(numpy.concatenate([numpy.arange(length - 1) % max(self.tokenizer.vocab_size - 1, 1) + 1, [self.tokenizer.eod]]))
        return sample

    def get(self, idx: int, offset: int = 0, length: Optional[int] = None) -> numpy.ndarray:
        if length is None:
            length = int(self.sequence_lengths[idx]) - offset
        return self[idx][offset : offset + length]


class MockGPTDataset(GPTDataset):
    def __init__(self, dataset: MockGPTLowLevelDataset, dataset_path: Optional[str], indices: numpy.ndarray, num_samples: int, index_split: Split,
                 config: GPTDatasetConfig):
        assert config.mock
        super().__init__(dataset, dataset_path, indices, num_samples, index_split, config)

    @staticmethod
    def numel_low_level_dataset(low_level_dataset: MockGPTLowLevelDataset) -> int:
        return len(low_level_dataset)

    @staticmethod
    def build_low_level_dataset(dataset_path: Optional[str], config: GPTDatasetConfig) -> MockGPTLowLevelDataset:
        return MockGPTLowLevelDataset(config.tokenizer)
