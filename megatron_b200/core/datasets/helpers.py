"""Thin typed front for the native index builders (reference ``datasets/helpers.py:11-66``)."""
from __future__ import annotations

import numpy


def _cpp():
    try:
        from . import helpers_cpp  # type: ignore
    except ImportError:
        from .utils import compile_helpers

        compile_helpers()
        from . import helpers_cpp  # type: ignore
    return helpers_cpp


def build_sample_idx(sizes: numpy.ndarray, document_indices: numpy.ndarray, sequence_length: int, num_epochs: int, tokens_per_epoch: int,
                     drop_last_partial_sequence: bool = True, add_extra_token_to_sequence: bool = True) -> numpy.ndarray:
    h = _cpp()
    sizes = numpy.ascontiguousarray(sizes, dtype=numpy.int32)
    document_indices = numpy.ascontiguousarray(document_indices, dtype=numpy.int32)
    big = max(len(document_indices), int(sizes.max()) if len(sizes) else 0) > numpy.iinfo(numpy.int32).max
    fn = h.build_sample_idx_int64 if big else h.build_sample_idx_int32
    return fn(sizes, document_indices, sequence_length, num_epochs, tokens_per_epoch, drop_last_partial_sequence, 1 if add_extra_token_to_sequence else 0)


def build_blending_indices(dataset_index, dataset_sample_index, weights, num_datasets, size, verbose=False):
    return _cpp().build_blending_indices(dataset_index, dataset_sample_index, numpy.asarray(weights, dtype=numpy.float64), num_datasets, size, verbose)


def build_exhaustive_blending_indices(dataset_index, dataset_sample_index, sizes, num_datasets):
    return _cpp().build_exhaustive_blending_indices(dataset_index, dataset_sample_index, numpy.asarray(sizes, dtype=numpy.int64), num_datasets)


def build_mapping(*a):
    return _cpp().build_mapping(*a)


def build_blocks_mapping(*a):
    return _cpp().build_blocks_mapping(*a)
