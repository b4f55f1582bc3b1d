"""Options of the torch-FSDP2 wrapper (reference ``distributed/torch_fully_sharded_data_parallel_config.py``)."""
from dataclasses import dataclass
from typing import Union

from .distributed_data_parallel_config import DistributedDataParallelConfig


@dataclass
class TorchFullyShardedDataParallelConfig(DistributedDataParallelConfig):
    reshard_after_forward: Union[bool, int] = True      # True: free the gathered parameters after forward; int: reshard to that smaller world size
