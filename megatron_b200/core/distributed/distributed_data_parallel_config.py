"""DDP configuration (reference ``distributed/distributed_data_parallel_config.py:15-243``)."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class DistributedDataParallelConfig:
    grad_reduce_in_fp32: bool = False
    overlap_grad_reduce: bool = False
    overlap_param_gather: bool = False
    align_param_gather: bool = False
    use_distributed_optimizer: bool = False
    num_distributed_optimizer_instances: int = 1
    check_for_nan_in_grad: bool = False
    check_for_large_grads: bool = False
    bucket_size: Optional[int] = None
    pad_buckets_for_high_nccl_busbw: bool = False
    average_in_collective: bool = False
    fp8_param_gather: bool = False
    fp4_param_gather: bool = False
    use_custom_fsdp: bool = False
    use_megatron_fsdp: bool = False
    data_parallel_sharding_strategy: str = "no_shard"
    gradient_reduce_div_fusion: bool = True
    suggested_communication_unit_size: Optional[int] = None
    preserve_fp32_weights: bool = True
    keep_fp8_transpose_cache: bool = False
    nccl_ub: bool = False
    fsdp_double_buffer: bool = False
    reduce_scatter_with_fp32_accumulation: bool = False
    delay_wgrad_compute: bool = False

    def __post_init__(self):
        if self.overlap_param_gather and not self.use_distributed_optimizer:
            raise ValueError("overlap_param_gather requires use_distributed_optimizer")
        if self.num_distributed_optimizer_instances > 1 and not self.use_distributed_optimizer:
            raise ValueError("num_distributed_optimizer_instances > 1 requires use_distributed_optimizer")
        if self.data_parallel_sharding_strategy not in ("no_shard", "optim", "optim_grads", "optim_grads_params"):
            raise ValueError(f"unknown sharding strategy {self.data_parallel_sharding_strategy}")
