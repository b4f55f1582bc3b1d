"""``.bin`` / ``.idx`` token store (reference ``datasets/indexed_dataset.py``).

Index file layout (little endian), kept byte-compatible with the reference so existing
preprocessed corpora load unchanged:
  9 B magic ``MMIDIDX\\x00\\x00`` | u64 version=1 | u8 dtype code | u64 n_sequences | u64 n_documents |
  i32[n_seq] lengths | i64[n_seq] byte pointers | i64[n_docs] document boundaries | (optional i8[n_seq] modes)
"""
from __future__ import annotations

import os
import struct
from enum import Enum
from typing import List, Optional, Union

import numpy
import torch

_INDEX_HEADER = b"MMIDIDX\x00\x00"


class DType(Enum):
    uint8 = 1
    int8 = 2
    int16 = 3
    int32 = 4
    int64 = 5
    float64 = 6
    float32 = 7
    uint16 = 8

    @classmethod
    def code_from_dtype(cls, value) -> int:
        return cls[numpy.dtype(value).name].value

    @classmethod
    def dtype_from_code(cls, code: int):
        return getattr(numpy, cls(code).name)

    @staticmethod
    def size(key) -> int:
        return numpy.dtype(DType.dtype_from_code(key) if isinstance(key, int) else key).itemsize

    @staticmethod
    def optimal_dtype(cardinality: Optional[int]):
        return numpy.uint16 if cardinality is not None and cardinality < 65500 else numpy.int32


def get_idx_path(prefix: str) -> str:
    return prefix + ".idx"


def get_bin_path(prefix: str) -> str:
    return prefix + ".bin"


class _IndexWriter:
    def __init__(self, idx_path: str, dtype):
        self.idx_path, self.dtype = idx_path, dtype

    def __enter__(self):
        self.f = open(self.idx_path, "wb")
        self.f.write(_INDEX_HEADER)
        self.f.write(struct.pack("<Q", 1))
        self.f.write(struct.pack("<B", DType.code_from_dtype(self.dtype)))
        return self

    def __exit__(self, *exc):
        self.f.close()

    def write(self, sequence_lengths: List[int], sequence_modes: Optional[List[int]], document_indices: List[int]):
        n = len(sequence_lengths)
        self.f.write(struct.pack("<Q", n))
        self.f.write(struct.pack("<Q", len(document_indices)))
        lens = numpy.array(sequence_lengths, dtype=numpy.int32)
        self.f.write(lens.tobytes(order="C"))
        itemsize = numpy.dtype(self.dtype).itemsize
        ptrs = numpy.zeros(n, dtype=numpy.int64)
        if n > 1:
            ptrs[1:] = numpy.cumsum(lens[:-1].astype(numpy.int64) * itemsize)
        self.f.write(ptrs.tobytes(order="C"))
        self.f.write(numpy.array(document_indices, dtype=numpy.int64).tobytes(order="C"))
        if sequence_modes is not None:
            self.f.write(numpy.array(sequence_modes, dtype=numpy.int8).tobytes(order="C"))


class _IndexReader:
    def __init__(self, idx_path: str, multimodal: bool):
        with open(idx_path, "rb") as f:
            assert f.read(9) == _INDEX_HEADER, f"bad header in {idx_path}"
            (version,) = struct.unpack("<Q", f.read(8))
            assert version == 1
            (code,) = struct.unpack("<B", f.read(1))
            self.dtype = DType.dtype_from_code(code)
            self.dtype_size = DType.size(self.dtype)
            (self.sequence_count,) = struct.unpack("<Q", f.read(8))
            (self.document_count,) = struct.unpack("<Q", f.read(8))
            off = f.tell()
        self.mm = numpy.memmap(idx_path, mode="r", order="C")
        buf = memoryview(self.mm)
        self.sequence_lengths = numpy.frombuffer(buf, dtype=numpy.int32, count=self.sequence_count, offset=off)
        off += self.sequence_lengths.nbytes
        self.sequence_pointers = numpy.frombuffer(buf, dtype=numpy.int64, count=self.sequence_count, offset=off)
        off += self.sequence_pointers.nbytes
        self.document_indices = numpy.frombuffer(buf, dtype=numpy.int64, count=self.document_count, offset=off)
        off += self.document_indices.nbytes
        self.sequence_modes = numpy.frombuffer(buf, dtype=numpy.int8, count=self.sequence_count, offset=off) if multimodal else None

    def __len__(self):
        return self.sequence_count

    def __getitem__(self, i):
        return self.sequence_pointers[i], self.sequence_lengths[i], (self.sequence_modes[i] if self.sequence_modes is not None else None)


class IndexedDataset(torch.utils.data.Dataset):
    def __init__(self, path_prefix: str, multimodal: bool = False, mmap: bool = True, **_):
        super().__init__()
        self.path_prefix, self.multimodal, self.mmap = path_prefix, multimodal, mmap
        self.index = _IndexReader(get_idx_path(path_prefix), multimodal)
        self.bin = numpy.memmap(get_bin_path(path_prefix), mode="r", order="C") if mmap else None
        self._fd = None if mmap else open(get_bin_path(path_prefix), "rb")

    def __getstate__(self):
        return self.path_prefix, self.multimodal, self.mmap

    def __setstate__(self, s):
        self.__init__(*s)

    def __len__(self):
        return len(self.index)

    def _read(self, ptr: int, count: int) -> numpy.ndarray:
        if self.bin is not None:
            return numpy.frombuffer(self.bin, dtype=self.index.dtype, count=count, offset=ptr)
        self._fd.seek(ptr)
        return numpy.frombuffer(self._fd.read(count * self.index.dtype_size), dtype=self.index.dtype)

    def __getitem__(self, idx: Union[int, slice]):
        if isinstance(idx, (int, numpy.integer)):
            ptr, length, mode = self.index[idx]
            seq = self._read(int(ptr), int(length))
            return (seq, mode) if mode is not None else seq
        start, stop, step = idx.indices(len(self))
        assert step == 1, "slices into indexed datasets must be contiguous"
        lens = self.index.sequence_lengths[idx]
        flat = self._read(int(self.index.sequence_pointers[start]), int(lens.sum()))
        return numpy.split(flat, numpy.cumsum(lens)[:-1])

    def get(self, idx: int, offset: int = 0, length: Optional[int] = None) -> numpy.ndarray:
        """Tokens ``[offset, offset+length)`` of sequence ``idx`` without touching the rest."""
        ptr, seq_len, mode = self.index[idx]
        if length is None:
            length = int(seq_len) - offset
        seq = self._read(int(ptr) + offset * self.index.dtype_size, length)
        return (seq, mode) if mode is not None else seq

    @property
    def sequence_lengths(self):
        return self.index.sequence_lengths

    @property
    def document_indices(self):
        return self.index.document_indices

    @staticmethod
    def exists(path_prefix: str) -> bool:
        return os.path.exists(get_idx_path(path_prefix)) and os.path.exists(get_bin_path(path_prefix))


class IndexedDatasetBuilder:
    def __init__(self, bin_path: str, dtype=numpy.int32, multimodal: bool = False):
        self.data_file = open(bin_path, "wb")
        self.dtype, self.multimodal = dtype, multimodal
        self.sequence_lengths: List[int] = []
        self.document_indices: List[int] = [0]
        self.sequence_modes: Optional[List[int]] = [] if multimodal else None

    def add_item(self, tensor, mode: int = 0):
        arr = numpy.asarray(tensor.numpy() if isinstance(tensor, torch.Tensor) else tensor, dtype=self.dtype)
        self.data_file.write(arr.tobytes(order="C"))
        self.sequence_lengths.append(arr.size)
        if self.multimodal:
            self.sequence_modes.append(mode)

    def add_document(self, tensor, lengths: List[int], modes: Optional[List[int]] = None):
        arr = numpy.asarray(tensor.numpy() if isinstance(tensor, torch.Tensor) else tensor, dtype=self.dtype)
        self.data_file.write(arr.tobytes(order="C"))
        self.sequence_lengths.extend(lengths)
        self.document_indices.append(len(self.sequence_lengths))
        if self.multimodal:
            self.sequence_modes.extend(modes if modes is not None else [0] * len(lengths))

    def end_document(self):
        self.document_indices.append(len(self.sequence_lengths))

    def add_index(self, path_prefix: str):
        idx = _IndexReader(get_idx_path(path_prefix), self.multimodal)
        assert idx.dtype == self.dtype
        off = len(self.sequence_lengths)
        self.sequence_lengths.extend(idx.sequence_lengths.tolist())
        self.document_indices.extend((off + idx.document_indices)[1:].tolist())
        if self.multimodal:
            self.sequence_modes.extend(idx.sequence_modes.tolist())
        with open(get_bin_path(path_prefix), "rb") as f:
            import shutil

            shutil.copyfileobj(f, self.data_file)

    def finalize(self, idx_path: str):
        self.data_file.close()
        with _IndexWriter(idx_path, self.dtype) as w:
            w.write(self.sequence_lengths, self.sequence_modes, self.document_indices)
