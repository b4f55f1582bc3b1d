from .distributed_data_parallel import DistributedDataParallel
from .distributed_data_parallel_config import DistributedDataParallelConfig
from .finalize_model_grads import finalize_model_grads

__all__ = ["DistributedDataParallel", "DistributedDataParallelConfig", "finalize_model_grads"]
from .fsdp import FullyShardedDataParallel
from .torch_fully_sharded_data_parallel import TorchFullyShardedDataParallel

__all__ += ["FullyShardedDataParallel", "TorchFullyShardedDataParallel"]
