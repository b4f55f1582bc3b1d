"""``TransformerConfig`` — model hyper-parameters + derived values.

Field names mirror the reference dataclass (``megatron/core/transformer/
transformer_config.py:57``) so that existing launch configs map 1:1; the
validation in ``__post_init__`` is written from the documented semantics.
New B200-specific knobs are grouped at the end (``b200_*``).
"""
from __future__ import annotations

import math

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from ..enums import AttnBackend
from ..model_parallel_config import ModelParallelConfig
from ..utils import init_method_normal, scaled_init_method_normal


@dataclass
class TransformerConfig(ModelParallelConfig):
    # ---- architecture ---------------------------------------------------------
    num_layers: int = 0
    mtp_num_layers: Optional[int] = None
    mtp_loss_scaling_factor: Optional[float] = 0.1
    mtp_grad_scale_func: Optional[Callable] = None
    # experimental attention variants (reference transformer_config.py:301-311): 'dsa' | 'gdn' | 'gated_delta_net'
    experimental_attention_variant: Optional[str] = None
    experimental_attention_variant_loss_scale_func: Optional[Callable] = None
    dsa_indexer_loss_coeff: float = 0.0
    linear_attention_freq: Optional[Union[int, List[int]]] = None      # gdn variant: int N = one softmax layer after every N-1 linear ones, or a 0/1 list
    # gated delta net mixer geometry (reference ``--linear-*``); None = the mixer's constructor arguments
    linear_conv_kernel_dim: Optional[int] = None
    linear_key_head_dim: Optional[int] = None
    linear_value_head_dim: Optional[int] = None
    linear_num_key_heads: Optional[int] = None
    linear_num_value_heads: Optional[int] = None
    # DeepSeek sparse attention index branch (reference ``transformer_config.py:315-357``); None = the module defaults (4 heads x 32, top-64)
    dsa_indexer_n_heads: Optional[int] = None
    dsa_indexer_head_dim: Optional[int] = None
    dsa_indexer_topk: Optional[int] = None
    dsa_indexer_use_sparse_loss: bool = False
    dsa_indexer_topk_freq: int = 1                     # > 1: only every n-th layer owns an indexer, the layers in between reuse its top-k
    dsa_indexer_skip_topk_offset: int = 0
    dsa_indexer_rope_interleaved: bool = False         # MLA-style interleaved rotation of the rope slice
    dsa_indexer_rotate_activation: bool = True         # Hadamard rotation of q / k before scoring
    dsa_indexer_scoring_relu: bool = True              # ReLU on q.k before the head weighting
    dsa_indexer_k_norm_epsilon: Optional[float] = None
    dsa_indexer_k_norm_fp32: bool = False
    num_layers_in_first_pipeline_stage: Optional[int] = None
    num_layers_in_last_pipeline_stage: Optional[int] = None
    pipeline_model_parallel_layout: Optional[Union[str, list]] = None
    account_for_embedding_in_pipeline_split: bool = False
    account_for_loss_in_pipeline_split: bool = False
    hidden_size: int = 0
    num_attention_heads: int = 0
    attention_backend: AttnBackend = AttnBackend.auto
    softmax_scale: Optional[float] = None
    # maximal-update parametrisation (reference ``transformer_config.py:428-475``): widths change, hyper-parameters transfer.  With ``use_mup`` the hidden
    # init std shrinks by 1/sqrt(m), the output-projection init by another depth factor, attention scores are scaled by 1/d_head (``mup_attn_scale_power`` 1.0)
    # instead of 1/sqrt(d_head), embeddings are multiplied by ``mup_embedding_mult`` and logits by ``mup_output_mult`` (auto 1/m), m = hidden / base hidden.
    use_mup: bool = False
    mup_width_mult: float = 1.0
    mup_base_hidden_size: Optional[int] = None
    mup_embedding_mult: float = 1.0
    mup_output_mult: float = 1.0
    mup_base_head_dim: Optional[float] = None
    mup_attn_scale_power: float = 1.0
    softmax_type: str = "vanilla"
    num_query_groups: Optional[int] = None
    ffn_hidden_size: Optional[int] = None
    kv_channels: Optional[int] = None
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    fp32_residual_connection: bool = False
    apply_residual_connection_post_layernorm: bool = False
    layernorm_epsilon: float = 1e-5
    layernorm_zero_centered_gamma: bool = False
    add_bias_linear: bool = True
    add_qkv_bias: bool = False
    gated_linear_unit: bool = False
    activation_func: Callable = F.gelu
    activation_func_fp8_input_store: bool = False
    glu_linear_offset: float = 0.0
    activation_func_clamp_value: Optional[float] = None
    num_moe_experts: Optional[int] = None
    rotary_interleaved: bool = False
    window_size: Optional[Tuple[int, int]] = None
    window_attn_skip_freq: Optional[Union[int, List[int]]] = None
    normalization: str = "LayerNorm"
    qk_layernorm: bool = False
    qk_l2_norm: bool = False
    qk_clip: bool = False
    qk_clip_alpha: float = 0.5
    qk_clip_threshold: float = 100
    attention_output_gate: bool = False
    test_mode: bool = False
    calculate_per_token_loss: bool = False
    multi_latent_attention: bool = False
    no_rope_freq: Optional[Union[int, List[int]]] = None

    # ---- init -----------------------------------------------------------------
    init_method: Optional[Callable] = None
    output_layer_init_method: Optional[Callable] = None
    init_method_std: float = 0.02
    embedding_init_method: Optional[Callable] = None
    embedding_init_method_std: Optional[float] = None
    init_model_with_meta_device: bool = False

    # ---- precision ------------------------------------------------------------
    apply_query_key_layer_scaling: bool = False
    attention_softmax_in_fp32: bool = True
    disable_bf16_reduced_precision_matmul: bool = False

    # ---- fusions (each maps to one sm_100a kernel in megatron_b200.ops) -------
    bias_activation_fusion: bool = False
    masked_softmax_fusion: bool = False
    persist_layer_norm: bool = False
    memory_efficient_layer_norm: bool = False
    bias_dropout_fusion: bool = False
    apply_rope_fusion: bool = False
    quant_recipe: Optional[object] = None          # core.quantization.RecipeConfig or a YAML path: per-layer FP8 / MXFP8 / NVFP4 / bf16 by module-path globs
    enable_mhc_connections: bool = False          # manifold-constrained hyper-connections: n-wide residual stream (transformer/hyper_connection.py)
    mhc_num_residual_streams: int = 4
    mhc_sinkhorn_iterations: int = 20
    mhc_init_gating_factor: float = 0.01
    mhc_recompute_layer_num: Optional[int] = None
    fused_single_qkv_rope: bool = False
    fused_residual_rmsnorm: bool = False
    use_fused_weighted_squared_relu: bool = False

    # ---- activation recompute -------------------------------------------------
    recompute_granularity: Optional[str] = None
    recompute_method: Optional[str] = None
    recompute_num_layers: Optional[int] = None
    distribute_saved_activations: Optional[bool] = False
    recompute_modules: Optional[List[str]] = None

    # ---- fp8 / fp4 ------------------------------------------------------------
    fp8: Optional[str] = None
    fp8_recipe: Optional[str] = "delayed"
    fp8_param: bool = False
    fp8_margin: int = 0
    fp8_interval: int = 1
    fp8_amax_history_len: int = 1
    fp8_amax_compute_algo: str = "most_recent"
    fp8_wgrad: bool = True
    fp8_dot_product_attention: bool = False
    fp8_multi_head_attention: bool = False
    tp_only_amax_red: bool = False
    first_last_layers_bf16: bool = False
    num_layers_at_start_in_bf16: int = 1
    num_layers_at_end_in_bf16: int = 1
    fp4: Optional[str] = None
    fp4_recipe: Optional[str] = "nvfp4"
    fp4_param: bool = False

    # ---- MoE ------------------------------------------------------------------
    moe_shared_expert_intermediate_size: Optional[int] = None
    moe_shared_expert_gate: bool = False
    moe_shared_expert_overlap: bool = False
    moe_layer_freq: Union[int, List[int]] = 1
    moe_ffn_hidden_size: Optional[int] = None
    moe_router_load_balancing_type: Union[str, List[str]] = "aux_loss"
    moe_router_topk: int = 2
    moe_router_topk_limited_devices: Optional[int] = None
    moe_router_padding_for_fp8: Optional[bool] = False
    moe_router_num_groups: Optional[int] = None
    moe_router_group_topk: Optional[int] = None
    moe_router_pre_softmax: bool = False
    moe_router_topk_scaling_factor: Optional[float] = None
    moe_router_score_function: str = "softmax"
    moe_router_dtype: Optional[str] = None
    moe_router_enable_expert_bias: bool = False
    moe_router_bias_update_rate: float = 1e-3
    moe_router_force_load_balancing: bool = False
    moe_grouped_gemm: bool = False
    moe_use_legacy_grouped_gemm: bool = False
    moe_aux_loss_coeff: Union[float, List[float]] = 0.0
    moe_z_loss_coeff: Optional[float] = None
    moe_input_jitter_eps: Optional[float] = None
    moe_token_dropping: bool = False
    moe_token_dispatcher_type: str = "allgather"
    moe_enable_deepep: bool = False
    moe_flex_dispatcher_backend: str = "b200"
    moe_per_layer_logging: bool = False
    moe_enable_routing_replay: bool = False
    moe_paged_stash: bool = False
    moe_paged_stash_page_size: int = 64
    inference_moe_token_dispatcher_type: Optional[str] = None   # None | nccl | nvls (static-shape serving dispatchers)
    moe_expert_capacity_factor: Optional[float] = None
    moe_pad_expert_input_to_capacity: bool = False
    moe_token_drop_policy: str = "probs"
    moe_layer_recompute: bool = False
    moe_permute_fusion: bool = False
    moe_router_fusion: bool = False
    moe_apply_probs_on_input: bool = False
    moe_deepep_num_sms: Optional[int] = 20

    # ---- context parallel -----------------------------------------------------
    cp_comm_type: Optional[Union[str, List[str]]] = None

    # ---- CUDA graphs ----------------------------------------------------------
    enable_cuda_graph: bool = False
    cuda_graph_use_single_mempool: bool = False
    cuda_graph_retain_backward_graph: bool = False
    cuda_graph_warmup_steps: int = 3
    external_cuda_graph: bool = False
    cuda_graph_impl: str = "none"
    cuda_graph_scope: Optional[Union[str, List[str]]] = None

    # ---- misc -----------------------------------------------------------------
    clone_scatter_output_in_embedding: bool = True
    disable_parameter_transpose_cache: bool = False
    config_logger_dir: str = ""
    flash_decode: bool = False
    inference_rng_tracker: bool = False
    inference_sampling_seed: int = 42
    symmetric_ar_type: Optional[str] = None
    mrope_section: Optional[List[int]] = None
    is_hybrid_model: bool = False
    mamba_state_dim: int = 128
    mamba_head_dim: int = 64
    mamba_num_groups: int = 8
    mamba_num_heads: Optional[int] = None
    use_mamba_mem_eff_path: bool = True
    mlp_chunks_for_prefill: int = 1
    mlp_chunks_for_training: int = 1
    fused_residual_rmsnorm: bool = False  # attention residual add + pre-MLP RMSNorm in one kernel (ops.add_rms_norm); needs no dropout / bias
    heterogeneous_block_specs: bool = False
    hetereogenous_dist_checkpoint: bool = False
    transformer_impl: str = "b200"
    fine_grained_activation_offloading: bool = False
    offload_modules: Optional[List[str]] = field(default_factory=list)

    # ---- B200-native knobs ----------------------------------------------------
    b200_fused_tp_comm: bool = True  # in-kernel NVLink AG→GEMM / GEMM→RS when tp>1 on GPU
    b200_comm_sms: int = 16  # SMs reserved for NVLink copy/reduce CTAs inside fused kernels
    b200_gemm_backend: str = "auto"  # "tcgen05" | "cublas" | "auto"
    b200_main_grads_dtype: torch.dtype = torch.float32

    def __post_init__(self):
        super().__post_init__()
        if self.fp16 and self.bf16:
            raise ValueError("fp16 and bf16 are mutually exclusive")
        if self.num_layers <= 0 and not self.is_hybrid_model:
            raise ValueError("num_layers must be positive")
        if self.hidden_size <= 0 or self.num_attention_heads <= 0:
            raise ValueError("hidden_size and num_attention_heads must be positive")
        if self.num_attention_heads % self.tensor_model_parallel_size != 0:
            raise ValueError(
                f"num_attention_heads ({self.num_attention_heads}) must be divisible by "
                f"tensor_model_parallel_size ({self.tensor_model_parallel_size})"
            )
        if self.ffn_hidden_size is None:
            self.ffn_hidden_size = 4 * self.hidden_size
        if self.kv_channels is None:
            self.kv_channels = self.hidden_size // self.num_attention_heads
        if self.num_query_groups is None:
            self.num_query_groups = self.num_attention_heads
        if self.num_attention_heads % self.num_query_groups != 0:
            raise ValueError("num_attention_heads must be a multiple of num_query_groups")
        if self.num_query_groups % self.tensor_model_parallel_size != 0 and self.num_query_groups >= self.tensor_model_parallel_size:
            raise ValueError("num_query_groups must be divisible by tensor_model_parallel_size")
        if self.apply_query_key_layer_scaling:
            self.attention_softmax_in_fp32 = True
        if self.expert_model_parallel_size > 1 and self.num_moe_experts is None:
            raise ValueError("expert parallelism needs num_moe_experts")
        if self.num_moe_experts is not None:
            if self.num_moe_experts <= 0:
                raise ValueError("num_moe_experts must be positive")
            if self.num_moe_experts % self.expert_model_parallel_size != 0:
                raise ValueError("num_moe_experts must be divisible by expert_model_parallel_size")
            if self.moe_ffn_hidden_size is None:
                self.moe_ffn_hidden_size = self.ffn_hidden_size
            if self.moe_router_topk > self.num_moe_experts:
                raise ValueError("moe_router_topk cannot exceed num_moe_experts")
            if self.moe_expert_capacity_factor is not None:
                if self.moe_expert_capacity_factor < 0:
                    self.moe_expert_capacity_factor = None
                elif self.moe_router_load_balancing_type not in ("aux_loss", "seq_aux_loss", "global_aux_loss", "none"):
                    raise ValueError("capacity factor requires an aux-loss or 'none' balancing type")
            if self.moe_pad_expert_input_to_capacity and self.moe_expert_capacity_factor is None:
                raise ValueError("moe_pad_expert_input_to_capacity needs moe_expert_capacity_factor")
            if self.moe_router_num_groups is not None:
                if self.num_moe_experts % self.moe_router_num_groups != 0:
                    raise ValueError("num_moe_experts must be divisible by moe_router_num_groups")
                if self.moe_router_group_topk is None:
                    raise ValueError("group-limited routing needs moe_router_group_topk")
        if self.enable_mhc_connections:
            if self.mtp_num_layers:
                raise ValueError("enable_mhc_connections is not compatible with multi-token prediction")
            if self.recompute_granularity == "full":
                raise ValueError("enable_mhc_connections supports selective recompute only")
            if self.pipeline_model_parallel_size > 1:
                raise ValueError("enable_mhc_connections needs pipeline_model_parallel_size == 1 in this build (the n-wide stream does not cross stages yet)")
            if self.mhc_num_residual_streams < 1 or self.mhc_sinkhorn_iterations < 1:
                raise ValueError("mhc_num_residual_streams and mhc_sinkhorn_iterations must be positive")
        if self.recompute_granularity is not None:
            if self.recompute_granularity not in ("full", "selective"):
                raise ValueError("recompute_granularity must be 'full' or 'selective'")
            if self.recompute_granularity == "full":
                if self.recompute_method not in ("uniform", "block"):
                    raise ValueError("full recompute needs recompute_method uniform|block")
                if self.recompute_num_layers is None:
                    raise ValueError("full recompute needs recompute_num_layers")
            else:
                if self.recompute_modules is None:
                    self.recompute_modules = ["core_attn"]
        if self.recompute_modules is None:
            self.recompute_modules = []
        allowed = {"core_attn", "moe_act", "mlp_act", "layernorm", "mla_up_proj", "mlp", "moe", "shared_experts", "mhc"}
        bad = set(self.recompute_modules) - allowed
        if bad:
            raise ValueError(f"unknown recompute modules {sorted(bad)}; allowed {sorted(allowed)}")
        if self.distribute_saved_activations and self.sequence_parallel:
            raise ValueError("distribute_saved_activations is incompatible with sequence_parallel")
        vp = self.virtual_pipeline_model_parallel_size
        if self.pipeline_model_parallel_size > 1 and self.pipeline_model_parallel_layout is None and not self.is_hybrid_model:
            # (hybrid stacks split by their layer pattern: "|" separators allow uneven stages, and the stack validates the split itself)
            n = self.num_layers
            first, last = self.num_layers_in_first_pipeline_stage, self.num_layers_in_last_pipeline_stage
            mid_stages = self.pipeline_model_parallel_size - (first is not None) - (last is not None)
            n_mid = n - (first or 0) - (last or 0)
            if self.account_for_embedding_in_pipeline_split:
                n_mid += 1
            if self.account_for_loss_in_pipeline_split:
                n_mid += 1
            if mid_stages > 0 and n_mid % mid_stages != 0:
                raise ValueError(f"{n_mid} layers cannot be split evenly over {mid_stages} pipeline stages")
            if vp is not None and mid_stages > 0 and (n_mid // mid_stages) % vp != 0:
                raise ValueError("layers per pipeline stage must be divisible by the virtual pipeline size")
        if self.bias_activation_fusion and self.activation_func not in (F.gelu, F.silu) and getattr(
            self.activation_func, "__name__", ""
        ) not in ("quick_gelu", "squared_relu"):
            raise ValueError("bias_activation_fusion supports gelu / silu (SwiGLU) / quick_gelu")
        if self.use_mup:
            if self.mup_base_hidden_size is None:
                self.mup_base_hidden_size = self.hidden_size
            if self.mup_base_hidden_size <= 0:
                raise ValueError("mup_base_hidden_size must be positive")
            self.mup_width_mult = self.hidden_size / self.mup_base_hidden_size
            if self.softmax_scale is None:
                kv = self.kv_channels or self.hidden_size // self.num_attention_heads
                self.softmax_scale = (1.0 if self.mup_base_head_dim is None else self.mup_base_head_dim**0.5) / (kv**self.mup_attn_scale_power)
            if self.mup_output_mult == 1.0 and self.mup_width_mult != 1.0:
                self.mup_output_mult = 1.0 / self.mup_width_mult
            if self.init_method is not None or self.output_layer_init_method is not None:
                import warnings

                warnings.warn("use_mup is enabled but a custom init_method / output_layer_init_method is set: the muP initialisation assumptions may not hold", UserWarning)
        # the embedding init is fixed BEFORE the hidden init picks up the muP width factor: embeddings keep the base std
        if self.embedding_init_method is None:
            std = self.embedding_init_method_std or self.init_method_std
            self.embedding_init_method = init_method_normal(std) if (self.init_method is None or std != self.init_method_std) else self.init_method
        depth_mult = 2.0 if not self.is_hybrid_model else 1.0
        width = self.mup_width_mult if self.use_mup else 1.0
        if self.init_method is None:
            self.init_method = init_method_normal(self.init_method_std / math.sqrt(width))
        if self.output_layer_init_method is None:
            self.output_layer_init_method = scaled_init_method_normal(self.init_method_std / math.sqrt(width), max(self.num_layers, 1), multiplier=depth_mult)
        if self.cp_comm_type is not None and isinstance(self.cp_comm_type, list):
            if len(self.cp_comm_type) != self.num_layers:
                raise ValueError("cp_comm_type list length must equal num_layers")
        if self.fp8 is not None and self.fp8 not in ("e4m3", "hybrid"):
            raise ValueError("fp8 must be 'e4m3' or 'hybrid'")
        if self.window_size is not None and len(self.window_size) != 2:
            raise ValueError("window_size must be (left, right)")
        if self.softmax_scale is None and self.apply_query_key_layer_scaling is False:
            pass  # default 1/sqrt(d) is applied inside attention
        if self.moe_token_dispatcher_type not in ("allgather", "alltoall", "flex", "alltoall_seq"):
            raise ValueError(f"unknown moe_token_dispatcher_type {self.moe_token_dispatcher_type}")
        if self.b200_main_grads_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("b200_main_grads_dtype must be fp32 or bf16")

    # convenience ---------------------------------------------------------------
    @property
    def head_dim(self) -> int:
        return self.kv_channels

    @property
    def compute_dtype(self) -> torch.dtype:
        return torch.bfloat16 if self.bf16 else (torch.float16 if self.fp16 else self.params_dtype)


@dataclass
class MLATransformerConfig(TransformerConfig):
    """Multi-latent attention (DeepSeek-V2/V3) config (reference :3285)."""

    multi_latent_attention: bool = True
    q_lora_rank: Optional[int] = 512
    kv_lora_rank: int = 512
    qk_head_dim: int = 128
    qk_pos_emb_head_dim: int = 64
    v_head_dim: int = 128
    normalization: str = "RMSNorm"
    rope_type: str = "yarn"
    rotary_base: float = 10000
    rotary_percent: float = 1.0
    rotary_scaling_factor: float = 40
    original_max_position_embeddings: int = 4096
    beta_fast: float = 32
    beta_slow: float = 1
    mscale: float = 1.0
    mscale_all_dim: float = 0.0
    cache_mla_latents: bool = False

    def __post_init__(self):
        super().__post_init__()
        if self.multi_latent_attention and self.apply_rope_fusion and self.rope_type != "yarn":
            raise ValueError("fused MLA rope requires yarn rope_type")
