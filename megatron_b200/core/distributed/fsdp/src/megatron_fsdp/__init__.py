from .fully_shard import fully_shard, fully_shard_model, fully_shard_optimizer  # noqa: F401
from .megatron_fsdp import MegatronFSDP  # noqa: F401
from .mixed_precision import MixedPrecisionPolicy  # noqa: F401
from .package_info import __version__  # noqa: F401
from .param_and_grad_buffer import (  # noqa: F401
    AllGatherPipeline, Bucket, BucketingPolicy, BucketStatus, DataParallelBuffer, FixedPoolAllocator, GradReducePipeline, MaxPoolAllocator,
    ParamAndGradBuffer, ParameterGroup, PrefetchOrder, RotaryBucketAllocator, StorageResizeBasedBucketAllocator, TemporaryBucketAllocator,
)
from .utils import FSDPDistributedIndex  # noqa: F401
