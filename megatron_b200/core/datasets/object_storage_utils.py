"""Datasets on object storage (reference ``datasets/object_storage_utils.py`` + ``utils_s3.py``): ``s3://bucket/key`` and
``msc://profile/key`` prefixes.  The ``.idx`` file is small and read whole, so it is cached on local disk once per node; the
``.bin`` file is read by byte range on demand (``IndexedDataset`` asks for ``(offset, nbytes)`` per sequence)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Protocol, Tuple

S3_PREFIX = "s3://"
MSC_PREFIX = "msc://"


@dataclass
class ObjectStorageConfig:
    path_to_idx_cache: str
    bin_chunk_nbytes: int = 256 * 1024 * 1024


S3Config = ObjectStorageConfig       # older name


class S3Client(Protocol):
    def download_file(self, Bucket: str, Key: str, Filename: str) -> None: ...
    def get_object(self, Bucket: str, Key: str, Range: str) -> Dict[str, Any]: ...
    def head_object(self, Bucket: str, Key: str) -> Dict[str, Any]: ...


def is_object_storage_path(path: str) -> bool:
    return isinstance(path, str) and (path.startswith(S3_PREFIX) or path.startswith(MSC_PREFIX))


is_s3_path = is_object_storage_path


def parse_s3_path(path: str) -> Tuple[str, str]:
    """``s3://bucket/a/b.bin`` → (``bucket``, ``a/b.bin``)."""
    for pre in (S3_PREFIX, MSC_PREFIX):
        if path.startswith(pre):
            rest = path[len(pre):]
            bucket, _, key = rest.partition("/")
            if not bucket:
                raise ValueError(f"no bucket in {path}")
            return bucket, key
    raise ValueError(f"{path} is not an object-storage path")


def get_index_cache_path(idx_path: str, cfg: ObjectStorageConfig) -> str:
    bucket, key = parse_s3_path(idx_path)
    return os.path.join(cfg.path_to_idx_cache, bucket, key)


def _client(path: str, client: Optional[S3Client] = None):
    if client is not None:
        return client
    if path.startswith(MSC_PREFIX):
        from ..msc_utils import MultiStorageClientFeature

        MultiStorageClientFeature.enable()
        return MultiStorageClientFeature.import_package()
    try:
        import boto3
    except ImportError as e:
        raise ImportError("reading s3:// datasets needs boto3 (not installed); pass a client= object or mirror the dataset locally") from e
    return boto3.client("s3")


def object_exists(path: str, client: Optional[S3Client] = None) -> bool:
    bucket, key = parse_s3_path(path)
    try:
        _client(path, client).head_object(Bucket=bucket, Key=key)
        return True
    except Exception:
        return False


def cache_index_file(remote_idx_path: str, cfg: ObjectStorageConfig, client: Optional[S3Client] = None, rank: Optional[int] = None) -> str:
    """Download the ``.idx`` once per node (local rank 0 downloads; the others wait at the caller's barrier) → local path."""
    local = get_index_cache_path(remote_idx_path, cfg)
    if os.path.exists(local):
        return local
    if rank is None:
        rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        os.makedirs(os.path.dirname(local), exist_ok=True)
        bucket, key = parse_s3_path(remote_idx_path)
        tmp = local + ".tmp"
        _client(remote_idx_path, client).download_file(Bucket=bucket, Key=key, Filename=tmp)
        os.replace(tmp, local)
    return local


class ObjectStorageBinReader:
    """Byte-range reads of a remote ``.bin`` with one resident chunk (sequential training reads hit it ~always)."""

    def __init__(self, path: str, cfg: ObjectStorageConfig, client: Optional[S3Client] = None):
        self.bucket, self.key = parse_s3_path(path)
        self.client = _client(path, client)
        self.chunk = cfg.bin_chunk_nbytes
        self._start, self._buf = 0, b""
        self.requests = 0

    def read(self, offset: int, nbytes: int) -> bytes:
        if not (self._start <= offset and offset + nbytes <= self._start + len(self._buf)):
            start = (offset // self.chunk) * self.chunk if nbytes <= self.chunk else offset
            end = max(start + self.chunk, offset + nbytes)
            obj = self.client.get_object(Bucket=self.bucket, Key=self.key, Range=f"bytes={start}-{end - 1}")
            body = obj["Body"]
            self._buf = body.read() if hasattr(body, "read") else bytes(body)
            self._start = start
            self.requests += 1
        a = offset - self._start
        return self._buf[a : a + nbytes]
