"""Serving configuration (reference ``inference/config.py:16-560``): one dataclass from which the dynamic engine, its paged KV
cache and the CUDA-graph decode buckets are sized.

    cfg = InferenceConfig(buffer_size_gb=40, block_size_tokens=16, max_sequence_length=8192, num_cuda_graphs=8)
    engine = cfg.build_engine(model)

B200 sizing notes: with 180 GB per GPU the KV budget is normally what is left after the weights (an 8B bf16 model leaves
≈150 GB = 1.2 M tokens of Llama-3-8B KV), so ``buffer_size_gb`` is a cap rather than a scarce resource; the block size stays
small (16 tokens) because the paged-attention kernel reads whole blocks and small blocks waste less on short tails."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional, Tuple

import torch


class PrefixCachingEvictionPolicy(str, Enum):
    REF_ZERO = "ref_zero"        # a block is dropped as soon as no request references it (no reuse across requests in time)
    LRU = "lru"                  # unreferenced blocks park in an LRU list and are evicted only under memory pressure


class PrefixCachingCoordinatorPolicy(str, Enum):
    ROUND_ROBIN = "round_robin"
    LONGEST_PREFIX = "longest_prefix"     # route to the data-parallel engine that already holds the longest cached prefix
    FIRST_PREFIX_BLOCK = "first_prefix_block"


class KVCacheManagementMode(str, Enum):
    PERSIST = "persist"          # KV memory stays allocated while the engine is suspended (RL: during the training phase)
    OFFLOAD = "offload"          # paged out to (pinned / managed) host memory and restored on resume
    RECOMPUTE = "recompute"      # released; in-flight requests are re-prefilled on resume


class CudaGraphSizingDistribution(str, Enum):
    LINEAR = "linear"
    EXPONENTIAL = "exponential"  # buckets 1, 2, 4, ... max: few graphs, ≤2x padding
    MIXED = "mixed"              # powers of two up to 16, then linear steps


class AsyncScheduleMode(str, Enum):
    LEGACY = "legacy"            # schedule -> launch -> wait, one step at a time
    OVERLAP = "overlap"          # step N+1 is scheduled on the host while step N runs on the device


@dataclass
class MambaInferenceStateConfig:
    """Shapes of the recurrent state a hybrid (Mamba / attention) model keeps per request."""
    layer_type_list: List[str]
    conv_states_shape: Tuple[int, ...]
    ssm_states_shape: Tuple[int, ...]
    conv_states_dtype: torch.dtype = torch.bfloat16
    ssm_states_dtype: torch.dtype = torch.float32
    mamba_chunk_size: int = 128
    ssm_chunk_alignment: Optional[int] = None
    gdp_num_householder: int = 0

    def bytes_per_request(self) -> int:
        n = sum(1 for t in self.layer_type_list if t in ("M", "G", "mamba"))
        conv = math.prod(self.conv_states_shape) * torch.empty((), dtype=self.conv_states_dtype).element_size()
        ssm = math.prod(self.ssm_states_shape) * torch.empty((), dtype=self.ssm_states_dtype).element_size()
        return n * (conv + ssm)

    @classmethod
    def from_model(cls, model, conv_states_dtype=None, ssm_states_dtype=None) -> Optional["MambaInferenceStateConfig"]:
        """Read the state geometry off the first Mamba mixer of a hybrid model; ``None`` for pure-attention models."""
        types, conv, ssm = [], None, None
        for m in model.modules():
            if hasattr(m, "mixer") or hasattr(m, "self_attention") or hasattr(m, "mlp"):
                mixer = getattr(m, "mixer", None)
                if mixer is not None and hasattr(mixer, "d_state"):
                    types.append("M")
                    if conv is None:
                        d_conv = getattr(mixer, "d_conv", 4)
                        conv_dim = getattr(mixer, "conv_dim", getattr(mixer, "d_inner", 0))
                        conv = (conv_dim, d_conv)
                        ssm = (getattr(mixer, "nheads", 1), getattr(mixer, "headdim", 1), mixer.d_state)
                elif hasattr(m, "self_attention"):
                    types.append("*")
        if conv is None:
            return None
        return cls(types, conv, ssm, conv_states_dtype or torch.bfloat16, ssm_states_dtype or torch.float32)


@dataclass
class ImageProcessingConfig:
    patch_dim: int
    dynamic_resolution: bool = False
    use_tiling: bool = False
    pixel_shuffle: bool = False
    spatial_merge_size: int = 1
    dynamic_resolution_min_patches: int = 1
    dynamic_resolution_max_patches: int = 128
    vision_model_type: str = "radio"
    pixel_mean: Optional[List[float]] = None
    pixel_std: Optional[List[float]] = None
    img_h: Optional[int] = None
    img_w: Optional[int] = None
    max_num_tiles: int = 1
    use_thumbnail: bool = False
    num_img_embeddings_per_tile: int = 0

    def embeddings_for(self, h: int, w: int) -> int:
        """Number of image-token embeddings an ``h x w`` image turns into."""
        ph, pw = h // self.patch_dim, w // self.patch_dim
        n = ph * pw
        if self.pixel_shuffle:
            n //= 4
        if self.spatial_merge_size > 1:
            n //= self.spatial_merge_size ** 2
        if self.dynamic_resolution:
            n = max(self.dynamic_resolution_min_patches, min(n, self.dynamic_resolution_max_patches))
        return n


@dataclass
class InferenceConfig:
    block_size_tokens: int = 16
    buffer_size_gb: float = 20
    paused_buffer_size_gb: Optional[float] = None
    mamba_inference_state_config: Optional[MambaInferenceStateConfig] = None
    mamba_memory_ratio: Optional[float] = None
    max_requests: Optional[int] = None
    max_tokens: Optional[int] = None
    unified_memory_level: int = 0
    kv_cache_management_mode: KVCacheManagementMode = KVCacheManagementMode.PERSIST
    num_cuda_graphs: Optional[int] = None
    cuda_graph_mixed_prefill_count: Optional[int] = 16
    cuda_graph_sizing_distribution: CudaGraphSizingDistribution = CudaGraphSizingDistribution.EXPONENTIAL
    use_cuda_graphs_for_non_decode_steps: bool = True
    cuda_graph_all_prefills: bool = False
    cuda_graph_max_tokens: int = 512
    static_kv_memory_pointers: bool = False
    max_sequence_length: int = 2560
    pg_collection: Optional[object] = None
    image_preprocessing_config: Optional[ImageProcessingConfig] = None
    use_flashinfer_fused_rope: Optional[bool] = False
    materialize_only_last_token_logits: bool = True
    enable_chunked_prefill: bool = False
    num_speculative_tokens: int = 0
    enable_prefix_caching: bool = False
    prefix_caching_eviction_policy: PrefixCachingEvictionPolicy = PrefixCachingEvictionPolicy.LRU
    prefix_caching_coordinator_policy: PrefixCachingCoordinatorPolicy = PrefixCachingCoordinatorPolicy.LONGEST_PREFIX
    prefix_caching_routing_alpha: float = 0.5
    prefix_caching_mamba_gb: Optional[float] = None
    track_paused_request_events: bool = False
    track_generated_token_events: bool = False
    metrics_writer: Optional[object] = None
    logging_step_interval: int = 0
    sampling_backend: str = "torch"
    offset_sampling_seed_by_dp_rank: bool = True
    async_sched_mode: AsyncScheduleMode = AsyncScheduleMode.LEGACY
    logprobs_mode: str = "raw_logprobs"
    request_metadata_types: Optional[List[Tuple[str, torch.dtype]]] = None
    use_synchronous_zmq_collectives: bool = False
    disable_ep_consensus: bool = False
    ep_consensus_interval: int = 20
    verbose: bool = False

    def __post_init__(self):
        for name, enum in (("kv_cache_management_mode", KVCacheManagementMode), ("cuda_graph_sizing_distribution", CudaGraphSizingDistribution),
                           ("prefix_caching_eviction_policy", PrefixCachingEvictionPolicy), ("prefix_caching_coordinator_policy", PrefixCachingCoordinatorPolicy),
                           ("async_sched_mode", AsyncScheduleMode)):
            setattr(self, name, enum(getattr(self, name)))
        if self.block_size_tokens <= 0 or self.block_size_tokens & (self.block_size_tokens - 1):
            raise ValueError(f"block_size_tokens must be a power of two, got {self.block_size_tokens}")
        if self.sampling_backend not in ("torch", "flashinfer"):
            raise ValueError(f"sampling_backend '{self.sampling_backend}'")
        if self.logprobs_mode not in ("raw_logprobs", "processed_logprobs"):
            raise ValueError(f"logprobs_mode '{self.logprobs_mode}'")
        if self.mamba_memory_ratio is not None and not 0.0 < self.mamba_memory_ratio < 1.0:
            raise ValueError("mamba_memory_ratio must be in (0, 1)")
        if self.num_speculative_tokens < 0:
            raise ValueError("num_speculative_tokens must be >= 0")

    # ---- sizing ------------------------------------------------------------------------------------------------------
    @staticmethod
    def kv_bytes_per_token(model_config, tp_size: int = 1, dtype: torch.dtype = torch.bfloat16) -> int:
        c = model_config
        layers = c.num_layers
        if getattr(c, "multi_latent_attention", False):
            width = c.kv_lora_rank + c.qk_pos_emb_head_dim          # the latent + the shared rope key, not per head, not TP-split
            return layers * width * torch.empty((), dtype=dtype).element_size()
        kv_heads = (getattr(c, "num_query_groups", None) or c.num_attention_heads)
        kv_heads = max(1, kv_heads // tp_size)
        head_dim = getattr(c, "kv_channels", None) or c.hidden_size // c.num_attention_heads
        return 2 * layers * kv_heads * head_dim * torch.empty((), dtype=dtype).element_size()

    def num_blocks(self, model_config, tp_size: int = 1, dtype: torch.dtype = torch.bfloat16) -> int:
        budget = self.buffer_size_gb * (1 << 30)
        if self.mamba_inference_state_config is not None and self.mamba_memory_ratio:
            budget *= 1.0 - self.mamba_memory_ratio
        per_block = self.kv_bytes_per_token(model_config, tp_size, dtype) * self.block_size_tokens
        n = int(budget // per_block)
        if self.max_tokens is not None:
            n = min(n, -(-self.max_tokens // self.block_size_tokens) + 1)
        return max(n, 2)

    def resolved_max_requests(self, model_config, tp_size: int = 1, dtype: torch.dtype = torch.bfloat16) -> int:
        if self.max_requests is not None:
            return self.max_requests
        # every running request needs at least one block; hybrid models are additionally bounded by the recurrent-state pool
        n = self.num_blocks(model_config, tp_size, dtype) - 1
        m = self.mamba_inference_state_config
        if m is not None and self.mamba_memory_ratio:
            n = min(n, int(self.buffer_size_gb * (1 << 30) * self.mamba_memory_ratio // max(1, m.bytes_per_request())))
        return max(1, min(n, 4096))

    def cuda_graph_batch_sizes(self, max_requests: int) -> List[int]:
        """Decode batch buckets a graph is captured for (ascending, always ending at ``max_requests``)."""
        if not self.num_cuda_graphs:
            return []
        n, top = self.num_cuda_graphs, max_requests
        if self.cuda_graph_sizing_distribution is CudaGraphSizingDistribution.LINEAR:
            step = max(1, -(-top // n))
            sizes = list(range(step, top + 1, step))
        elif self.cuda_graph_sizing_distribution is CudaGraphSizingDistribution.EXPONENTIAL:
            sizes = [1 << i for i in range(top.bit_length()) if (1 << i) <= top]
            sizes = sizes[-n:] if len(sizes) > n else sizes
            if n > 1 and sizes[0] != 1 and len(sizes) == n:
                sizes[0] = 1                                          # always keep the single-request graph (interactive latency)
        else:
            small = [b for b in (1, 2, 4, 8, 16) if b <= top]
            rest = max(0, n - len(small))
            step = max(1, -(-(top - 16) // rest)) if rest and top > 16 else 0
            sizes = small + (list(range(16 + step, top + 1, step)) if step else [])
        sizes = sorted(set(s for s in sizes if 0 < s <= top))
        if not sizes or sizes[-1] != top:
            sizes.append(top)
        return sizes

    def build_engine(self, model, vocab_size: Optional[int] = None, tp_size: int = 1):
        """The dynamic engine (paged KV cache, chunked prefill, prefix caching, graphed decode) this configuration describes."""
        from .engine import DynamicInferenceEngine
        mc = model.config
        dtype = getattr(mc, "params_dtype", torch.bfloat16)
        max_req = self.resolved_max_requests(mc, tp_size, dtype)
        buckets = self.cuda_graph_batch_sizes(max_req)
        return DynamicInferenceEngine(
            model, num_blocks=self.num_blocks(mc, tp_size, dtype), block_size=self.block_size_tokens, max_running=max_req, vocab_size=vocab_size,
            enable_prefix_caching=self.enable_prefix_caching, decode_batch_buckets=buckets or None, enable_cuda_graphs=bool(buckets) and torch.cuda.is_available(),
            max_prefill_tokens_per_step=(self.max_tokens if self.enable_chunked_prefill else None))
