"""Optimizer configuration (reference ``optimizer/optimizer_config.py``)."""
from dataclasses import dataclass
from typing import Callable, Optional

import torch


@dataclass
class OptimizerConfig:
    optimizer: str = "adam"
    lr: Optional[float] = None
    min_lr: Optional[float] = None
    decoupled_lr: Optional[float] = None
    decoupled_min_lr: Optional[float] = None
    weight_decay: float = 0.01
    fp16: bool = False
    bf16: bool = False
    params_dtype: torch.dtype = torch.float32
    use_precision_aware_optimizer: bool = False
    main_grads_dtype: torch.dtype = torch.float32
    main_params_dtype: torch.dtype = torch.float32
    exp_avg_dtype: torch.dtype = torch.float32
    exp_avg_sq_dtype: torch.dtype = torch.float32
    loss_scale: Optional[float] = None
    initial_loss_scale: float = 2**32
    min_loss_scale: float = 1.0
    loss_scale_window: float = 1000
    hysteresis: int = 2
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-08
    decoupled_weight_decay: bool = True
    sgd_momentum: float = 0.9
    # Muon (emerging optimizers): orthogonalised momentum for 2-D weights, AdamW for everything else
    muon_momentum: float = 0.95
    muon_nesterov: bool = True
    muon_ns_steps: int = 5
    muon_scale_mode: str = "spectral"        # spectral: sqrt(max(1, out/in)) | shape: 0.2 * sqrt(max(out, in)) (match-AdamW-RMS) | none
    muon_tp_mode: str = "blockwise"          # blockwise: orthogonalise each TP shard | duplicated: gather the full matrix over TP first
    muon_extra_scale: float = 1.0
    soap_shampoo_beta: float = 0.95          # EMA of the Kronecker statistics G Gᵀ / Gᵀ G
    soap_precondition_frequency: int = 10    # steps between eigenbasis refreshes
    soap_max_precond_dim: int = 8192         # a side larger than this keeps the identity basis (one-sided SOAP)
    soap_precondition_warmup: bool = True    # the first step only seeds the statistics
    qk_clip_threshold: Optional[float] = None  # MuonClip: cap on the max attention logit (None = off)
    qk_clip_alpha: float = 0.5
    use_mup: bool = False                      # maximal-update parametrisation: hidden-matrix lr scaled by base_hidden / hidden
    mup_base_hidden_size: Optional[int] = None
    optimizer_cpu_offload: bool = False
    optimizer_offload_fraction: float = 1.0
    overlap_cpu_optimizer_d2h_h2d: bool = False
    use_distributed_optimizer: bool = False
    overlap_param_gather: bool = False
    overlap_param_gather_with_optimizer_step: bool = False
    optimizer_cpu_offload: bool = False
    optimizer_offload_fraction: float = 0.0
    clip_grad: float = 1.0
    log_num_zeros_in_grad: bool = False
    barrier_with_L1_time: bool = False
    timers: Optional[Callable] = None
    config_logger_dir: str = ""

    def __post_init__(self):
        if self.fp16 and self.bf16:
            raise ValueError("fp16 and bf16 are mutually exclusive")
        if self.optimizer not in ("adam", "sgd", "lion", "muon", "soap"):
            raise ValueError(f"unknown optimizer {self.optimizer}")
