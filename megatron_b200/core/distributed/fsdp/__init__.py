from .fully_sharded_data_parallel import FullyShardedDataParallel  # noqa: F401
