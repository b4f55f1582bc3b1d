"""High-level training engine — the public "one call" API.

    engine = TrainEngine("llama3_8b", tensor_model_parallel_size=8, sequence_parallel=True,
                         micro_batch_size=1, global_batch_size=4)
    loss = engine.train_step(tokens_cpu)      # tokens: pinned int64 [global_batch/dp, seq+1]

It wires together what the reference's ``megatron/training/training.py`` does in
``setup_model_and_optimizer`` (:2665) + ``train_step`` (:3010): model chunks → DDP buffers →
(distributed) optimizer → schedule → finalize grads → fused optimizer step → LR schedule.
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, List, Optional

import torch
import torch.distributed as dist

from ..core import parallel_state as ps
from ..core.distributed import DistributedDataParallel, DistributedDataParallelConfig, finalize_model_grads
from ..core.optimizer import OptimizerConfig, get_megatron_optimizer
from ..core.optimizer_param_scheduler import OptimizerParamScheduler
from ..core.pipeline_parallel.schedules import get_forward_backward_func
from ..core.tensor_parallel.random import model_parallel_cuda_manual_seed
from ..models.presets import build_gpt_model
from .flops import num_floating_point_operations


def _device():
    if torch.cuda.is_available() and (not dist.is_initialized() or dist.get_backend() != "gloo"):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def initialize_distributed(backend: Optional[str] = None):
    """Process-group bootstrap from torchrun env vars (single process works too)."""
    if dist.is_initialized():
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kw = {}
    if backend == "nccl":
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)


class TrainEngine:
    def __init__(self, model: str = "tiny_llama", *, tensor_model_parallel_size: int = 1, pipeline_model_parallel_size: int = 1,
                 virtual_pipeline_model_parallel_size: Optional[int] = None, context_parallel_size: int = 1, expert_model_parallel_size: int = 1,
                 sequence_parallel: Optional[bool] = None, micro_batch_size: int = 1, global_batch_size: int = 1, seq_length: Optional[int] = None,
                 bf16: bool = True, lr: float = 3e-4, min_lr: float = 3e-5, weight_decay: float = 0.1, clip_grad: float = 1.0,
                 lr_warmup_samples: int = 0, lr_decay_samples: Optional[int] = None, use_distributed_optimizer: bool = True,
                 overlap_grad_reduce: bool = True, overlap_param_gather: bool = False, grad_reduce_in_fp32: bool = False, seed: int = 1234,
                 recompute_granularity: Optional[str] = None, recompute_modules: Optional[List[str]] = None, recompute_method=None,
                 recompute_num_layers=None, nvlink_collectives: Optional[bool] = None, gradient_accumulation_fusion: bool = True,
                 model_overrides: Optional[dict] = None, ddp_bucket_size: Optional[int] = None):
        initialize_distributed()
        self.device = _device()
        if not ps.is_initialized():
            ps.initialize_model_parallel(
                tensor_model_parallel_size=tensor_model_parallel_size, pipeline_model_parallel_size=pipeline_model_parallel_size,
                virtual_pipeline_model_parallel_size=virtual_pipeline_model_parallel_size, context_parallel_size=context_parallel_size,
                expert_model_parallel_size=expert_model_parallel_size, create_gloo_process_groups=False,
            )
        tp, pp = tensor_model_parallel_size, pipeline_model_parallel_size
        if sequence_parallel is None:
            sequence_parallel = tp > 1
        model_parallel_cuda_manual_seed(seed)
        dtype = torch.bfloat16 if bf16 else torch.float32
        ov = dict(
            tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp, virtual_pipeline_model_parallel_size=virtual_pipeline_model_parallel_size,
            context_parallel_size=context_parallel_size, expert_model_parallel_size=expert_model_parallel_size,
            sequence_parallel=sequence_parallel, bf16=bf16, params_dtype=dtype, pipeline_dtype=dtype if pp > 1 else None,
            use_cpu_initialization=self.device.type == "cpu", gradient_accumulation_fusion=gradient_accumulation_fusion,
            recompute_granularity=recompute_granularity, recompute_modules=recompute_modules, recompute_method=recompute_method,
            recompute_num_layers=recompute_num_layers,
        )
        if seq_length is not None:
            ov["seq_length"] = seq_length
        if os.environ.get("MEGATRON_B200_FUSED_RESIDUAL_NORM") is not None:
            ov["fused_residual_rmsnorm"] = os.environ["MEGATRON_B200_FUSED_RESIDUAL_NORM"] == "1"
        ov.update(model_overrides or {})
        # NVLink (symmetric-memory) collectives for the TP group when on GPUs of one box
        if nvlink_collectives is None:
            from ..parallel import fused as _fused

            nvlink_collectives = self.device.type == "cuda" and tp > 1 and _fused.get_mode(world_size=tp) != "nccl"
        if nvlink_collectives:
            from ..parallel import collectives
            from ..parallel import fused as _fused

            try:
                # workspace = two halves, each holding the largest fused pair-op footprint: a [seq*mbs, hidden] bf16 partial-sum / gathered buffer PLUS the
                # piggy-backed all-gather of the wgrad operand of the same size (Llama-3 70B: 2 x 128 MiB per half)
                from ..models.presets import PRESETS

                pr = PRESETS.get(model, {})
                rows = (seq_length or pr.get("seq_length", 0)) * micro_batch_size
                hid = (model_overrides or {}).get("hidden_size", pr.get("hidden_size", 0))
                need = 4 * rows * hid * 2 + (8 << 20)
                env_mb = os.environ.get("MEGATRON_B200_NVL_WORKSPACE_MB")
                ws_bytes = (int(env_mb) << 20) if env_mb else max(320 << 20, need)
                collectives.enable_for_group(ps.get_tensor_model_parallel_group(), workspace_bytes=ws_bytes)
            except Exception as e:  # no symmetric memory / NVLS on this box: keep training over NCCL (all ranks fail alike)
                import warnings

                warnings.warn(f"NVLink symmetric-memory runtime unavailable ({type(e).__name__}: {e}); TP collectives fall back to NCCL")
                _fused.set_mode("nccl")

        # data-parallel gradient reduce-scatter / parameter all-gather over NVLink: the DDP buffers are then allocated from this group's symmetric heap
        # (MEGATRON_B200_DP_COMM=nccl keeps torch.distributed)
        if self.device.type == "cuda" and os.environ.get("MEGATRON_B200_DP_COMM", "nvlink") != "nccl":
            dp_group = ps.get_data_parallel_group(with_context_parallel=True)
            if dist.get_world_size(dp_group) > 1 and expert_model_parallel_size == 1:
                from ..parallel import collectives

                try:
                    collectives.enable_for_group(dp_group, workspace_bytes=64 << 20)
                except Exception as e:
                    import warnings

                    warnings.warn(f"NVLink symmetric-memory runtime unavailable for the DP group ({type(e).__name__}: {e}); gradients reduce over NCCL")
        # expert-parallel NVLink dispatch/combine ("flex" dispatcher) needs a symmetric heap on the EP group
        if self.device.type == "cuda" and expert_model_parallel_size > 1 and ov.get("moe_token_dispatcher_type") == "flex":
            from ..parallel import collectives

            collectives.enable_for_group(ps.get_expert_model_parallel_group())

        vp = virtual_pipeline_model_parallel_size
        self.model_chunks = []
        for v in range(vp or 1):
            if vp is not None:
                ps.set_virtual_pipeline_model_parallel_rank(v)
            pre = ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=v if vp else None)
            post = ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=v if vp else None)
            m, cfg, preset = build_gpt_model(model, pre_process=pre, post_process=post, vp_stage=v if vp else None, **dict(ov))
            self.model_chunks.append(m.to(self.device))
        if vp is not None:
            ps.set_virtual_pipeline_model_parallel_rank(0)
        self.config, self.preset = cfg, preset
        self.seq_length = preset["seq_length"]
        self.micro_batch_size, self.global_batch_size = micro_batch_size, global_batch_size
        dp = ps.get_data_parallel_world_size()
        assert global_batch_size % (micro_batch_size * dp) == 0, "global batch must be divisible by micro_batch * dp"
        self.num_microbatches = global_batch_size // (micro_batch_size * dp)

        ddp_cfg = DistributedDataParallelConfig(
            grad_reduce_in_fp32=grad_reduce_in_fp32, overlap_grad_reduce=overlap_grad_reduce and dp > 1,
            overlap_param_gather=overlap_param_gather and dp > 1, use_distributed_optimizer=use_distributed_optimizer, bucket_size=ddp_bucket_size,
        )
        self.model = [DistributedDataParallel(c.config, ddp_cfg, c) for c in self.model_chunks]
        # all model chunks share ONE config object; wire the callbacks the schedules use
        for c in self.model_chunks:
            c.config = self.config
        opt_cfg = OptimizerConfig(optimizer="adam", lr=lr, min_lr=min_lr, weight_decay=weight_decay, bf16=bf16, params_dtype=dtype,
                                  clip_grad=clip_grad, use_distributed_optimizer=use_distributed_optimizer, adam_beta1=0.9, adam_beta2=0.95)
        self.optimizer = get_megatron_optimizer(opt_cfg, self.model)
        decay = lr_decay_samples or (global_batch_size * 100000)
        self.scheduler = OptimizerParamScheduler(
            self.optimizer, init_lr=0.0 if lr_warmup_samples else lr, max_lr=lr, min_lr=min_lr, lr_warmup_steps=lr_warmup_samples, lr_decay_steps=decay,
            lr_decay_style="cosine", start_wd=weight_decay, end_wd=weight_decay, wd_incr_steps=decay, wd_incr_style="constant",
            use_checkpoint_opt_param_scheduler=False,
        )
        self.config.finalize_model_grads_func = finalize_model_grads
        self.config.grad_scale_func = self.optimizer.scale_loss if (self.config.fp16) else None
        if len(self.model) == 1:
            self.config.no_sync_func = self.model[0].no_sync
        else:
            self.config.no_sync_func = [m.no_sync for m in self.model]
        self.forward_backward_func = get_forward_backward_func()
        self.iteration = 0
        self._position_ids = None
        self.flops_per_step = num_floating_point_operations(
            **{k: preset[k] for k in ("num_layers", "hidden_size", "ffn_hidden_size", "num_attention_heads", "num_query_groups", "kv_channels", "vocab_size")},
            seq_length=self.seq_length, batch_size=global_batch_size, swiglu=preset["swiglu"],
            num_moe_experts=preset.get("num_moe_experts"), moe_router_topk=preset.get("moe_router_topk", 1),
        )

    # ---- data ----------------------------------------------------------------------------------
    def _microbatch_iter(self, tokens_dev: torch.Tensor) -> Iterator[Dict[str, torch.Tensor]]:
        mbs = self.micro_batch_size
        s = self.seq_length
        if self._position_ids is None or self._position_ids.shape != (mbs, s):
            self._position_ids = torch.arange(s, device=tokens_dev.device).unsqueeze(0).expand(mbs, s).contiguous()
        for i in range(self.num_microbatches):
            chunk = tokens_dev[i * mbs : (i + 1) * mbs]
            yield {"tokens": chunk[:, :-1].contiguous(), "labels": chunk[:, 1:].contiguous(), "position_ids": self._position_ids}

    @staticmethod
    def _loss_func(output_tensor: torch.Tensor):
        loss = output_tensor.float().mean()
        return loss, {"lm loss": loss.detach()}

    def _forward_step(self, data_iterator, model):
        from ..core.utils import get_batch_on_this_cp_rank

        b = next(data_iterator)
        if self.config.context_parallel_size > 1:
            b = get_batch_on_this_cp_rank(b)
        out = model(b["tokens"], b["position_ids"], None, labels=b["labels"])
        return out, self._loss_func

    # ---- step ------------------------------------------------------------------------------------
    def train_step(self, tokens: torch.Tensor) -> torch.Tensor:
        """One optimizer step.  ``tokens``: int64 ``[global_batch/dp, seq+1]`` on the host (pinned)
        or already on the device.  Returns the mean loss (0-d tensor on the device; only meaningful
        on the last pipeline stage)."""
        tokens_dev = tokens.to(self.device, non_blocking=True) if tokens.device != self.device else tokens
        for m in self.model:
            m.zero_grad_buffer()
        self.optimizer.zero_grad()
        its = [self._microbatch_iter(tokens_dev) for _ in self.model]
        losses = self.forward_backward_func(
            forward_step_func=self._forward_step, data_iterator=its if len(its) > 1 else its[0], model=self.model if len(self.model) > 1 else self.model[0],
            num_microbatches=self.num_microbatches, seq_length=self.seq_length, micro_batch_size=self.micro_batch_size, forward_only=False,
        )
        ok, grad_norm, _ = self.optimizer.step()
        self.last_grad_norm = grad_norm
        if ok:
            self.scheduler.step(self.global_batch_size)
        self.iteration += 1
        if losses:
            return torch.stack([d["lm loss"] for d in losses]).mean()
        return torch.zeros((), device=self.device)

    @torch.no_grad()
    def eval_step(self, tokens: torch.Tensor) -> torch.Tensor:
        tokens_dev = tokens.to(self.device, non_blocking=True)
        for m in self.model:
            m.eval()
        its = [self._microbatch_iter(tokens_dev) for _ in self.model]
        losses = self.forward_backward_func(
            forward_step_func=self._forward_step, data_iterator=its if len(its) > 1 else its[0], model=self.model if len(self.model) > 1 else self.model[0],
            num_microbatches=self.num_microbatches, seq_length=self.seq_length, micro_batch_size=self.micro_batch_size, forward_only=True,
        )
        for m in self.model:
            m.train()
        return torch.stack([d["lm loss"] for d in losses]).mean() if losses else torch.zeros((), device=self.device)

    def synthetic_batch(self, pinned: bool = True, seed: int = 0) -> torch.Tensor:
        """Random token ids of the benchmark's shape on the host."""
        g = torch.Generator().manual_seed(seed)
        n = self.num_microbatches * self.micro_batch_size
        t = torch.randint(0, self.preset["vocab_size"], (n, self.seq_length + 1), generator=g, dtype=torch.int64)
        return t.pin_memory() if (pinned and torch.cuda.is_available()) else t
