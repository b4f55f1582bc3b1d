"""Distributed (sharded, parallelism-independent) checkpointing."""
from .core import check_is_distributed_checkpoint
from .mapping import LocalNonpersistentObject, ShardedObject, ShardedTensor
from .serialization import load, load_common_state_dict, load_plain_tensors, load_tensors_metadata, save
from .validation import StrictHandling
from .serialization import load_content_metadata, remove_sharded_tensors  # noqa: E402,F401
