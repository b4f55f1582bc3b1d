"""Tensor/sequence-parallel linear layers and vocab-parallel embedding.

API parity with reference ``tensor_parallel/layers.py`` (``VocabParallelEmbedding``
:232, ``ColumnParallelLinear`` :932, ``RowParallelLinear`` :1328, TP parameter
attributes :138-168).  The compute/communication core is different: every
linear goes through one of two *pair ops* from ``megatron_b200.parallel.fused``

* ``all_gather_gemm``     — sequence-parallel all-gather + ``X Wᵀ``
* ``gemm_reduce_scatter`` — ``X Wᵀ`` + reduce-scatter (or all-reduce)

which on B200 are single sm_100a kernels that move tiles over NVLink while
tcgen05 computes, and on CPU/Gloo decompose into a collective and a matmul.
Backward uses the dual pair ops (dgrad GEMM→RS, AG→wgrad GEMM with fp32
``main_grad`` accumulation in the epilogue).
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional

import torch
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from .. import parallel_state as ps
from ..model_parallel_config import ModelParallelConfig
from ..utils import divide, get_pg_rank, get_pg_size, get_tensor_model_parallel_group_if_none
from .mappings import (
    copy_to_tensor_model_parallel_region,
    gather_from_sequence_parallel_region,
    gather_from_tensor_model_parallel_region,
    reduce_from_tensor_model_parallel_region,
    reduce_scatter_to_sequence_parallel_region,
    scatter_to_tensor_model_parallel_region,
)
from .random import get_cuda_rng_tracker, get_expert_parallel_rng_tracker_name
from .utils import VocabUtility

_MODEL_PARALLEL_ATTRIBUTE_DEFAULTS = {"tensor_model_parallel": False, "partition_dim": -1, "partition_stride": 1}


def param_is_not_tensor_parallel_duplicate(param, tp_group=None) -> bool:
    """True for TP-sharded params, and for replicated params on TP rank 0."""
    if getattr(param, "tensor_model_parallel", False):
        return True
    if tp_group is not None:
        return get_pg_rank(tp_group) == 0
    return ps.get_tensor_model_parallel_rank() == 0


def set_tensor_model_parallel_attributes(tensor, is_parallel, dim, stride):
    for a in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        assert not hasattr(tensor, a)
    tensor.tensor_model_parallel = is_parallel
    tensor.partition_dim = dim
    tensor.partition_stride = stride


def set_defaults_if_not_set_tensor_model_parallel_attributes(tensor):
    for a, v in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS.items():
        if not hasattr(tensor, a):
            setattr(tensor, a, v)


def copy_tensor_model_parallel_attributes(dst, src):
    for a in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        if hasattr(src, a):
            setattr(dst, a, getattr(src, a))


def _init_sharded_weight(weight, init_method, partition_dim, stride=1, is_expert=False):
    """Initialise a shard in place with the per-TP-rank RNG stream."""
    set_tensor_model_parallel_attributes(weight, True, partition_dim, stride)
    tracker = get_cuda_rng_tracker()
    if not tracker.is_initialized():
        init_method(weight)
        return
    name = get_expert_parallel_rng_tracker_name() if is_expert else "model-parallel-rng"
    with tracker.fork(name):
        init_method(weight)


def _init_sharded_weight_from_master(weight, out_size, in_size, per_partition_size, partition_dim, init_method, stride, rank, world_size, params_dtype, return_master=False):
    """Initialise the *full* matrix deterministically on CPU, keep this rank's slice."""
    set_tensor_model_parallel_attributes(weight, True, partition_dim, stride)
    master = torch.empty(out_size, in_size, dtype=torch.float, requires_grad=False)
    init_method(master)
    master = master.to(dtype=params_dtype)
    per_stride = divide(per_partition_size, stride)
    pieces = torch.split(master, per_stride, dim=partition_dim)
    mine = pieces[rank::world_size]
    with torch.no_grad():
        weight.data.copy_(torch.cat(mine, dim=partition_dim))
    return master if return_master else None


# kept under the reference names for users that import them
_initialize_affine_weight_gpu = _init_sharded_weight
_initialize_affine_weight_cpu = _init_sharded_weight_from_master


def _device_for(config):
    if config.use_cpu_initialization or not torch.cuda.is_available():
        return "cpu"
    return torch.cuda.current_device()


class VocabParallelEmbedding(torch.nn.Module):
    """Embedding table sharded along the vocabulary axis."""

    def __init__(
        self,
        num_embeddings: int,
        embedding_dim: int,
        *,
        init_method: Callable,
        reduce_scatter_embeddings: bool = False,
        config: ModelParallelConfig,
        tp_group=None,
        pg_collection=None,
    ):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.reduce_scatter_embeddings = reduce_scatter_embeddings
        self.tp_group = get_tensor_model_parallel_group_if_none(tp_group)
        ws, rk = get_pg_size(self.tp_group), get_pg_rank(self.tp_group)
        self.vocab_start_index, self.vocab_end_index = VocabUtility.vocab_range_from_global_vocab_size(num_embeddings, rk, ws)
        self.num_embeddings_per_partition = self.vocab_end_index - self.vocab_start_index
        self.deterministic_mode = config.deterministic_mode
        self.tensor_model_parallel_size = ws
        dev = _device_for(config)
        self.weight = Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim, device=dev, dtype=config.params_dtype))
        if config.perform_initialization:
            if config.use_cpu_initialization:
                _init_sharded_weight_from_master(
                    self.weight, num_embeddings, embedding_dim, self.num_embeddings_per_partition, 0, init_method, 1, rk, ws, config.params_dtype
                )
            else:
                _init_sharded_weight(self.weight, init_method, partition_dim=0, stride=1)
        else:
            set_tensor_model_parallel_attributes(self.weight, True, 0, 1)

    def forward(self, input_):
        if self.tensor_model_parallel_size > 1:
            mask = (input_ < self.vocab_start_index) | (input_ >= self.vocab_end_index)
            local = (input_ - self.vocab_start_index).masked_fill(mask, 0)
        else:
            local, mask = input_, None
        out = F.embedding(local, self.weight) if not self.deterministic_mode else self.weight[local]
        if mask is not None:
            out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        if self.reduce_scatter_embeddings:
            # [b, s, h] → [s, b, h] then RS along the sequence
            out = out.transpose(0, 1).contiguous()
            return reduce_scatter_to_sequence_parallel_region(out, group=self.tp_group)
        return reduce_from_tensor_model_parallel_region(out, group=self.tp_group)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..utils import make_tp_sharded_tensor_for_checkpoint

        sd = self.state_dict(prefix="", keep_vars=True)
        key = f"{prefix}weight"
        return {
            key: make_tp_sharded_tensor_for_checkpoint(
                sd["weight"], key, allow_shape_mismatch=True, prepend_offsets=sharded_offsets, tp_group=self.tp_group
            )
        }


class _TPLinearFn(torch.autograd.Function):
    """y = X Wᵀ with the TP/SP collectives folded into the GEMM pair ops.

    Parity: ``LinearWithGradAccumulationAndAsyncCommunication`` reference :580-810.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, grad_accum_fusion, allreduce_dgrad, sequence_parallel, tp_group, wgrad_deferral_limit, grad_output_buffer):
        from ...parallel import fused

        ctx.use_bias = bias is not None
        ctx.grad_accum_fusion = grad_accum_fusion
        ctx.allreduce_dgrad = allreduce_dgrad
        ctx.sequence_parallel = sequence_parallel
        ctx.tp_group = tp_group
        ctx.grad_output_buffer = grad_output_buffer
        ctx.wgrad_deferral_limit = wgrad_deferral_limit
        ctx.save_for_backward(x, weight)
        if sequence_parallel:
            out = fused.all_gather_gemm(x, weight, tp_group)
        else:
            out = fused.gemm_nt(x, weight)
        if bias is not None:
            out = out + bias
        return out

    @staticmethod
    def backward(ctx, gy):
        from ...parallel import fused

        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        wgrad_needed = weight.requires_grad
        if ctx.grad_output_buffer is not None and wgrad_needed:
            ctx.grad_output_buffer.append(gy)
            wgrad_needed = False
        gx = gw = gb = None
        if ctx.sequence_parallel:
            # dgrad GEMM → reduce-scatter; all-gather(X) → wgrad GEMM.  The fused
            # module overlaps the two pairs on separate streams / in-kernel.
            gx, gw = fused.sp_linear_backward(gy, x, weight, ctx.tp_group, wgrad_needed, ctx.grad_accum_fusion)
        else:
            handle = None
            if ctx.allreduce_dgrad and get_pg_size(ctx.tp_group) > 1:
                gx, handle = fused.dgrad_all_reduce(gy, weight, ctx.tp_group)      # fused GEMM->AR kernel, or GEMM + async collective
            else:
                gx = fused.gemm_nn(gy, weight)
            if wgrad_needed:
                gw = fused.wgrad(gy, x, weight, ctx.grad_accum_fusion)
            if handle is not None:
                handle.wait()
        if ctx.use_bias:
            gb = gy.reshape(-1, gy.shape[-1]).sum(dim=0)
        return gx, gw, gb, None, None, None, None, None, None


def linear_with_grad_accumulation_and_async_allreduce(
    input, weight, bias, gradient_accumulation_fusion, allreduce_dgrad, sequence_parallel,
    grad_output_buffer=None, wgrad_deferral_limit=0, async_grad_allreduce=None, tp_group=None,
):
    if async_grad_allreduce is not None:
        warnings.warn("async_grad_allreduce is deprecated; use allreduce_dgrad")
        allreduce_dgrad = async_grad_allreduce
    tp_group = get_tensor_model_parallel_group_if_none(tp_group)
    return _TPLinearFn.apply(
        input, weight, bias, gradient_accumulation_fusion, allreduce_dgrad, sequence_parallel, tp_group, wgrad_deferral_limit, grad_output_buffer
    )


class _FrozenLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, allreduce_dgrad, tp_group):
        ctx.save_for_backward(weight)
        ctx.allreduce_dgrad, ctx.tp_group = allreduce_dgrad, tp_group
        out = torch.matmul(x, weight.t())
        return out + bias if bias is not None else out

    @staticmethod
    def backward(ctx, gy):
        (weight,) = ctx.saved_tensors
        gx = gy.matmul(weight)
        if ctx.allreduce_dgrad and get_pg_size(ctx.tp_group) > 1:
            torch.distributed.all_reduce(gx, group=ctx.tp_group)
        return gx, None, None, None, None


def linear_with_frozen_weight(input, weight, bias, gradient_accumulation_fusion, allreduce_dgrad, sequence_parallel, tp_group=None, **_):
    tp_group = get_tensor_model_parallel_group_if_none(tp_group)
    if sequence_parallel:
        input = gather_from_sequence_parallel_region(input, tensor_parallel_output_grad=True, group=tp_group)
    return _FrozenLinearFn.apply(input, weight, bias, allreduce_dgrad, tp_group)


def _fp8_active(config) -> bool:
    if not getattr(config, "fp8", None):
        return False
    from ..fp8_utils import fp8_enabled

    return fp8_enabled()


def _fp8_recipe(config) -> str:
    r = getattr(config, "fp8_recipe", None) or "tensorwise"
    return {"delayed": "tensorwise", "blockwise": "mxfp8"}.get(r, r)      # delayed scaling needs per-layer meta objects: falls back to current scaling here


def _fp8_choice(module):
    """→ ``(recipe, fp8_format)`` for THIS layer or ``None`` (bf16).  A per-layer ``quant_config`` (attached by ``core.quantization.apply_quantization_recipe``
    from glob matchers over module paths) wins over the model-wide ``config.fp8`` / autocast state."""
    qc = getattr(module, "quant_config", None)
    if qc is not None:
        if not qc.enabled:
            return None
        return {"delayed": "tensorwise", "blockwise": "mxfp8"}.get(qc.recipe, qc.recipe), qc.fp8_format
    cfg = module.config
    return (_fp8_recipe(cfg), cfg.fp8) if _fp8_active(cfg) else None


class ColumnParallelLinear(torch.nn.Module):
    """Y = X Aᵀ with A sharded along its output dimension."""

    def __init__(
        self,
        input_size,
        output_size,
        *,
        config: ModelParallelConfig,
        init_method: Callable,
        bias=True,
        gather_output=False,
        stride=1,
        keep_master_weight_for_test=False,
        skip_bias_add=False,
        skip_weight_param_allocation: bool = False,
        embedding_activation_buffer: Optional[List[torch.Tensor]] = None,
        grad_output_buffer: Optional[List[torch.Tensor]] = None,
        is_expert: bool = False,
        tp_comm_buffer_name: str = None,
        disable_grad_reduce: bool = False,
        tp_group=None,
        name=None,
        output_dtype=None,
        pg_collection=None,
    ):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.gather_output, self.skip_bias_add, self.is_expert = gather_output, skip_bias_add, is_expert
        self.config = config
        self.embedding_activation_buffer = embedding_activation_buffer
        self.grad_output_buffer = grad_output_buffer
        self.disable_grad_reduce = disable_grad_reduce
        self.tp_group = get_tensor_model_parallel_group_if_none(tp_group, is_expert=is_expert)
        ws, rk = get_pg_size(self.tp_group), get_pg_rank(self.tp_group)
        self.explicit_expert_comm = is_expert and (ws > 1 or getattr(config, "expert_model_parallel_size", 1) > 1)
        self.output_size_per_partition = divide(output_size, ws)
        self.tp_comm_buffer_name = tp_comm_buffer_name
        dev = _device_for(config)
        self.master_weight = None
        if not skip_weight_param_allocation:
            self.weight = Parameter(torch.empty(self.output_size_per_partition, input_size, device=dev, dtype=config.params_dtype))
            if config.perform_initialization:
                if config.use_cpu_initialization:
                    self.master_weight = _init_sharded_weight_from_master(
                        self.weight, output_size, input_size, self.output_size_per_partition, 0, init_method, stride, rk, ws,
                        config.params_dtype, return_master=keep_master_weight_for_test,
                    )
                else:
                    _init_sharded_weight(self.weight, init_method, partition_dim=0, stride=stride, is_expert=is_expert)
            else:
                set_tensor_model_parallel_attributes(self.weight, True, 0, stride)
            setattr(self.weight, "allreduce", not (is_expert and getattr(config, "expert_model_parallel_size", 1) > 1))
        else:
            self.weight = None
        if bias:
            self.bias = Parameter(torch.zeros(self.output_size_per_partition, device=dev, dtype=config.params_dtype))
            set_tensor_model_parallel_attributes(self.bias, True, 0, stride)
            setattr(self.bias, "allreduce", not (is_expert and getattr(config, "expert_model_parallel_size", 1) > 1))
        else:
            self.register_parameter("bias", None)
        self.sequence_parallel = config.sequence_parallel and ws > 1
        self.allreduce_dgrad = ws > 1 and not self.sequence_parallel and not disable_grad_reduce
        self.gradient_accumulation_fusion = config.gradient_accumulation_fusion
        if self.explicit_expert_comm:
            self.sequence_parallel = self.allreduce_dgrad = False
        self._register_load_state_dict_pre_hook(
            lambda sd, prefix, *a: sd.setdefault(f"{prefix}_extra_state", None) and None
        )

    def forward(self, input_, weight=None, runtime_gather_output=None):
        if weight is None:
            if self.weight is None:
                raise RuntimeError("weight was not allocated and none was passed to forward")
            weight = self.weight
        else:
            exp = (self.output_size_per_partition, self.input_size)
            if tuple(weight.shape) != exp:
                raise RuntimeError(f"supplied weight has shape {tuple(weight.shape)}, expected {exp}")
        bias = self.bias if not self.skip_bias_add else None
        if self.allreduce_dgrad or self.sequence_parallel or self.explicit_expert_comm or self.disable_grad_reduce:
            x = input_
        else:
            x = copy_to_tensor_model_parallel_region(input_, group=self.tp_group)
        if self.config.defer_embedding_wgrad_compute and self.embedding_activation_buffer is not None:
            self.embedding_activation_buffer.append(x)
        fp8 = _fp8_choice(self)
        if fp8 is not None and not self.explicit_expert_comm:
            # FP8 / MXFP8 recipe: collectives through the autograd mappings, the three GEMMs of the layer through core.fp8_utils.fp8_linear
            from ..fp8_utils import fp8_linear

            if self.sequence_parallel and fp8[0] == "mxfp8" and getattr(self.config, "fp8_quantized_all_gather", True) and x.shape[-1] % 128 == 0 \
                    and get_pg_size(self.tp_group) > 1:
                from ..fp8_utils import mxfp8_sp_column_linear

                out_parallel = mxfp8_sp_column_linear(x, weight, self.tp_group)        # activation shard goes on the wire as MXFP8 (half the bytes)
            else:
                xg = gather_from_sequence_parallel_region(x, tensor_parallel_output_grad=True, group=self.tp_group) if self.sequence_parallel else (
                    copy_to_tensor_model_parallel_region(x, group=self.tp_group) if self.allreduce_dgrad else x)
                out_parallel = fp8_linear(xg, weight, recipe=fp8[0], fp8_format=fp8[1])
            if bias is not None:
                out_parallel = out_parallel + bias
            gather = self.gather_output if runtime_gather_output is None else runtime_gather_output
            out = gather_from_tensor_model_parallel_region(out_parallel, group=self.tp_group) if gather else out_parallel
            return out, (self.bias if self.skip_bias_add else None)
        fn = linear_with_grad_accumulation_and_async_allreduce if weight.requires_grad else linear_with_frozen_weight
        out_parallel = fn(
            input=x, weight=weight, bias=bias,
            gradient_accumulation_fusion=self.gradient_accumulation_fusion,
            allreduce_dgrad=False if self.explicit_expert_comm else self.allreduce_dgrad,
            sequence_parallel=False if self.explicit_expert_comm else self.sequence_parallel,
            grad_output_buffer=self.grad_output_buffer if self.config.defer_embedding_wgrad_compute else None,
            wgrad_deferral_limit=self.config.wgrad_deferral_limit if self.config.defer_embedding_wgrad_compute else 0,
            tp_group=self.tp_group,
        )
        gather = self.gather_output if runtime_gather_output is None else runtime_gather_output
        if gather:
            assert not self.sequence_parallel
            out = gather_from_tensor_model_parallel_region(out_parallel, group=self.tp_group)
        else:
            out = out_parallel
        return out, (self.bias if self.skip_bias_add else None)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..transformer.utils import make_sharded_tensors_for_checkpoint

        sd = self.state_dict(prefix="", keep_vars=True)
        return make_sharded_tensors_for_checkpoint(sd, prefix, {"weight": 0, "bias": 0}, sharded_offsets, tp_group=self.tp_group)

    def set_extra_state(self, state):
        pass

    def get_extra_state(self):
        return None

    def __repr__(self):
        return f"{type(self).__name__}(in_features={self.input_size}, out_features={self.output_size}, bias={self.bias is not None}, TP={get_pg_size(self.tp_group)})"


class _RowLinearFn(torch.autograd.Function):
    """Row-parallel y = reduce(X_shard W_shardᵀ): GEMM → reduce-scatter (SP) or all-reduce."""

    @staticmethod
    def forward(ctx, x, weight, sequence_parallel, tp_group, grad_accum_fusion):
        from ...parallel import fused

        ctx.save_for_backward(x, weight)
        ctx.sequence_parallel, ctx.tp_group, ctx.grad_accum_fusion = sequence_parallel, tp_group, grad_accum_fusion
        if tp_group is None or get_pg_size(tp_group) == 1:
            return fused.gemm_nt(x, weight)
        if sequence_parallel:
            return fused.gemm_reduce_scatter(x, weight, tp_group)
        return fused.gemm_all_reduce(x, weight, tp_group)

    @staticmethod
    def backward(ctx, gy):
        from ...parallel import fused

        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        if ctx.sequence_parallel and ctx.tp_group is not None and get_pg_size(ctx.tp_group) > 1:
            # all-gather(dY) feeds both dgrad and wgrad
            gx, gw = fused.row_linear_backward_sp(gy, x, weight, ctx.tp_group, weight.requires_grad, ctx.grad_accum_fusion)
        else:
            gx = fused.gemm_nn(gy, weight)
            gw = fused.wgrad(gy, x, weight, ctx.grad_accum_fusion) if weight.requires_grad else None
        return gx, gw, None, None, None


class RowParallelLinear(torch.nn.Module):
    """Y = X Aᵀ with A sharded along its input dimension."""

    def __init__(
        self,
        input_size: int,
        output_size: int,
        *,
        config: ModelParallelConfig,
        init_method: Callable,
        bias: bool,
        input_is_parallel: bool,
        skip_bias_add: bool,
        stride: int = 1,
        keep_master_weight_for_test: bool = False,
        is_expert: bool = False,
        tp_comm_buffer_name: str = None,
        tp_group=None,
        name=None,
        pg_collection=None,
    ):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.input_is_parallel, self.skip_bias_add, self.is_expert = input_is_parallel, skip_bias_add, is_expert
        self.config = config
        self.tp_group = get_tensor_model_parallel_group_if_none(tp_group, is_expert=is_expert)
        ws, rk = get_pg_size(self.tp_group), get_pg_rank(self.tp_group)
        self.explicit_expert_comm = is_expert and (ws > 1 or getattr(config, "expert_model_parallel_size", 1) > 1)
        self.input_size_per_partition = divide(input_size, ws)
        self.sequence_parallel = config.sequence_parallel and ws > 1
        if self.sequence_parallel and not input_is_parallel:
            raise RuntimeError("sequence parallelism requires input_is_parallel=True")
        self.gradient_accumulation_fusion = config.gradient_accumulation_fusion
        dev = _device_for(config)
        self.master_weight = None
        self.weight = Parameter(torch.empty(output_size, self.input_size_per_partition, device=dev, dtype=config.params_dtype))
        if config.perform_initialization:
            if config.use_cpu_initialization:
                self.master_weight = _init_sharded_weight_from_master(
                    self.weight, output_size, input_size, self.input_size_per_partition, 1, init_method, stride, rk, ws,
                    config.params_dtype, return_master=keep_master_weight_for_test,
                )
            else:
                _init_sharded_weight(self.weight, init_method, partition_dim=1, stride=stride, is_expert=is_expert)
        else:
            set_tensor_model_parallel_attributes(self.weight, True, 1, stride)
        setattr(self.weight, "allreduce", not (is_expert and getattr(config, "expert_model_parallel_size", 1) > 1))
        if bias:
            self.bias = Parameter(torch.zeros(output_size, device=dev, dtype=config.params_dtype))
            setattr(self.bias, "allreduce", not (is_expert and getattr(config, "expert_model_parallel_size", 1) > 1))
            setattr(self.bias, "sequence_parallel", self.sequence_parallel)
        else:
            self.register_parameter("bias", None)
        self._register_load_state_dict_pre_hook(
            lambda sd, prefix, *a: sd.setdefault(f"{prefix}_extra_state", None) and None
        )

    def forward(self, input_):
        x = input_ if self.input_is_parallel else scatter_to_tensor_model_parallel_region(input_, group=self.tp_group)
        fp8 = _fp8_choice(self)
        if fp8 is not None and not self.explicit_expert_comm and self.weight.requires_grad:
            from ..fp8_utils import fp8_linear

            out = fp8_linear(x, self.weight, recipe=fp8[0], fp8_format=fp8[1])
            if get_pg_size(self.tp_group) > 1:
                out = (reduce_scatter_to_sequence_parallel_region if self.sequence_parallel else reduce_from_tensor_model_parallel_region)(out, group=self.tp_group)
        elif self.explicit_expert_comm:
            from ...parallel import fused

            out = _RowLinearFn.apply(x, self.weight, False, None, self.gradient_accumulation_fusion)
        elif self.weight.requires_grad:
            out = _RowLinearFn.apply(x, self.weight, self.sequence_parallel, self.tp_group, self.gradient_accumulation_fusion)
        else:
            out = torch.matmul(x, self.weight.t())
            out = (reduce_scatter_to_sequence_parallel_region if self.sequence_parallel else reduce_from_tensor_model_parallel_region)(out, group=self.tp_group)
        if not self.skip_bias_add:
            return (out + self.bias if self.bias is not None else out), None
        return out, self.bias

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..transformer.utils import make_sharded_tensors_for_checkpoint

        sd = self.state_dict(prefix="", keep_vars=True)
        return make_sharded_tensors_for_checkpoint(sd, prefix, {"weight": 1}, sharded_offsets, tp_group=self.tp_group)

    def set_extra_state(self, state):
        pass

    def get_extra_state(self):
        return None

    def __repr__(self):
        return f"{type(self).__name__}(in_features={self.input_size}, out_features={self.output_size}, bias={self.bias is not None}, TP={get_pg_size(self.tp_group)})"
