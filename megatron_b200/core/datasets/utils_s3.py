"""Kept for import parity (reference ``datasets/utils_s3.py`` re-exports ``object_storage_utils``)."""
from .object_storage_utils import *  # noqa: F401,F403
from .object_storage_utils import S3Config, is_s3_path, parse_s3_path  # noqa: F401
