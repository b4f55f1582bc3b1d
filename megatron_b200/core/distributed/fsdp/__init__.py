from .fully_sharded_data_parallel import FullyShardedDataParallel  # noqa: F401
# the layered implementation (bucket allocators, DataParallelBuffer, AllGather / GradReduce pipelines, fully_shard API)
from .src.megatron_fsdp import MegatronFSDP, fully_shard  # noqa: F401,E402
