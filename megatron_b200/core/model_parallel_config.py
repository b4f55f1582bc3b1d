"""``ModelParallelConfig`` — parallelism knobs shared by every module.

Field names follow the reference (``megatron/core/model_parallel_config.py:46-575``)
so user configs carry over.  Knobs that selected TransformerEngine userbuffer
overlap in the reference (``tp_comm_overlap*`` :265-340) select the in-kernel
NVLink fusion here: ``tp_comm_overlap=True`` routes Column/RowParallelLinear
through the fused all-gather→GEMM / GEMM→reduce-scatter kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, ContextManager, Optional

import torch


@dataclass
class ModelParallelConfig:
    # ---- model parallelism ----------------------------------------------------
    tensor_model_parallel_size: int = 1
    pipeline_model_parallel_comm_backend: Optional[str] = None
    pipeline_model_parallel_size: int = 1
    virtual_pipeline_model_parallel_size: Optional[int] = None
    sequence_parallel: bool = False
    context_parallel_size: int = 1
    hierarchical_context_parallel_sizes: Optional[list] = None
    hybrid_context_parallel: bool = False
    max_seqlen_per_dp_cp_rank: Optional[int] = None
    expert_model_parallel_size: int = 1
    expert_tensor_parallel_size: Optional[int] = None
    moe_extended_tp: bool = False

    # ---- initialisation -------------------------------------------------------
    perform_initialization: bool = True
    use_cpu_initialization: bool = False

    # ---- precision ------------------------------------------------------------
    fp16: bool = False
    bf16: bool = False
    params_dtype: torch.dtype = torch.float32
    timers: Optional[Callable] = None
    finalize_model_grads_func: Optional[Callable] = None
    grad_scale_func: Optional[Callable] = None
    no_sync_func: Optional[Callable] = None
    grad_sync_func: Optional[Callable] = None
    param_sync_func: Optional[Callable] = None
    deterministic_mode: bool = False
    enable_autocast: bool = False
    autocast_dtype: Optional[torch.dtype] = None
    num_microbatches_with_partial_activation_checkpoints: Optional[int] = None

    # ---- fusion / overlap -----------------------------------------------------
    gradient_accumulation_fusion: bool = False
    async_tensor_model_parallel_allreduce: bool = True
    use_te_rng_tracker: bool = False
    tp_comm_overlap: bool = False
    tp_comm_bulk_wgrad: bool = True
    tp_comm_bulk_dgrad: bool = True
    tp_comm_overlap_ag: bool = True
    tp_comm_overlap_rs: bool = True
    tp_comm_overlap_rs_dgrad: bool = False
    tp_comm_split_ag: bool = True
    tp_comm_atomic_ag: bool = False
    tp_comm_split_rs: bool = True
    tp_comm_atomic_rs: bool = False
    cross_entropy_loss_fusion: bool = False
    cross_entropy_fusion_impl: str = "native"
    tp_comm_overlap_disable_qkv: bool = False
    tp_comm_overlap_disable_fc1: bool = False
    tp_comm_bootstrap_backend: str = "nccl"
    overlap_moe_expert_parallel_comm: bool = False
    delay_wgrad_compute: bool = False

    # ---- pipeline -------------------------------------------------------------
    pipeline_dtype: Optional[torch.dtype] = None
    variable_seq_lengths: bool = False
    overlap_p2p_comm: bool = False
    batch_p2p_comm: bool = True
    batch_p2p_sync: bool = True
    use_ring_exchange_p2p: bool = False
    deallocate_pipeline_outputs: bool = False
    defer_embedding_wgrad_compute: bool = False
    wgrad_deferral_limit: int = 0
    overlap_p2p_comm_warmup_flush: bool = False
    microbatch_group_size_per_vp_stage: Optional[int] = None

    # ---- cpu offload ----------------------------------------------------------
    cpu_offloading: bool = False
    cpu_offloading_num_layers: int = 0
    _cpu_offloading_context: Optional[ContextManager] = None
    cpu_offloading_activations: bool = True
    cpu_offloading_weights: bool = False
    cpu_offloading_double_buffering: bool = False

    # ---- misc -----------------------------------------------------------------
    barrier_with_L1_time: bool = True

    def __post_init__(self):
        if self.sequence_parallel and self.tensor_model_parallel_size <= 1:
            raise ValueError("sequence_parallel requires tensor_model_parallel_size > 1")
        if self.expert_tensor_parallel_size is None:
            self.expert_tensor_parallel_size = self.tensor_model_parallel_size
        if self.pipeline_model_parallel_size > 1 and self.pipeline_dtype is None:
            raise ValueError("pipeline_dtype must be set when pipeline_model_parallel_size > 1")
        if self.autocast_dtype is None:
            self.autocast_dtype = self.params_dtype
        if self.defer_embedding_wgrad_compute and self.pipeline_model_parallel_size == 1:
            raise ValueError("defer_embedding_wgrad_compute needs pipeline parallelism")
        if self.defer_embedding_wgrad_compute and not self.gradient_accumulation_fusion:
            raise ValueError("defer_embedding_wgrad_compute needs gradient_accumulation_fusion")
        if self.defer_embedding_wgrad_compute and self.wgrad_deferral_limit < 0:
            raise ValueError("wgrad_deferral_limit must be >= 0")
        if self.virtual_pipeline_model_parallel_size is not None and self.virtual_pipeline_model_parallel_size > 1:
            if self.microbatch_group_size_per_vp_stage is None:
                self.microbatch_group_size_per_vp_stage = self.pipeline_model_parallel_size
            if self.overlap_p2p_comm and self.batch_p2p_comm:
                raise ValueError("overlap_p2p_comm requires batch_p2p_comm=False for interleaved schedules")
        if self.expert_model_parallel_size > 1 and self.tensor_model_parallel_size > 1 and not self.sequence_parallel:
            raise ValueError("tensor parallel + expert parallel requires sequence_parallel")
