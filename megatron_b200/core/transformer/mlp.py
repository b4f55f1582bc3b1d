"""Dense feed-forward block (reference ``transformer/mlp.py:155``).

fc1 (column-parallel, all-gather fused in) → activation (SwiGLU/GeGLU/GeLU
sm_100a kernels) → fc2 (row-parallel, reduce-scatter fused in).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch
import torch.nn.functional as F

from ... import ops
from ..dist_checkpointing.mapping import ReplicaId, ShardedStateDict, ShardedTensor, ShardedTensorFactory
from ..utils import get_tensor_model_parallel_group_if_none
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import TransformerConfig
from .utils import sharded_state_dict_default


@dataclass
class MLPSubmodules:
    linear_fc1: Union[ModuleSpec, type] = None
    linear_fc2: Union[ModuleSpec, type] = None
    activation_func: Union[ModuleSpec, type] = None


class MLP(MegatronModule):
    def __init__(self, config: TransformerConfig, submodules: MLPSubmodules, is_expert: bool = False,
                 input_size: Optional[int] = None, ffn_hidden_size: Optional[int] = None, tp_group=None):
        super().__init__(config)
        self.input_size = input_size if input_size is not None else config.hidden_size
        self.tp_group = get_tensor_model_parallel_group_if_none(tp_group, is_expert=is_expert)
        if ffn_hidden_size is None:
            ffn_hidden_size = config.moe_ffn_hidden_size if is_expert else config.ffn_hidden_size
        self.ffn_hidden_size = ffn_hidden_size
        fc1_out = ffn_hidden_size * (2 if config.gated_linear_unit else 1)
        self.linear_fc1 = build_module(
            submodules.linear_fc1, self.input_size, fc1_out, config=config, init_method=config.init_method,
            gather_output=False, bias=config.add_bias_linear, skip_bias_add=True, is_expert=is_expert,
            tp_comm_buffer_name="fc1", tp_group=self.tp_group,
        )
        self.activation_func = config.activation_func
        self.linear_fc2 = build_module(
            submodules.linear_fc2, ffn_hidden_size, config.hidden_size, config=config, init_method=config.output_layer_init_method,
            bias=config.add_bias_linear, input_is_parallel=True, skip_bias_add=True, is_expert=is_expert,
            tp_comm_buffer_name="fc2", tp_group=self.tp_group,
        )

    def _activation(self, y, bias, per_token_scale=None):
        cfg = self.config
        if cfg.gated_linear_unit:
            if self.activation_func is F.silu and cfg.activation_func_clamp_value is None and cfg.glu_linear_offset == 0.0:
                return ops.swiglu(y, bias, per_token_scale)
            if self.activation_func is F.gelu and per_token_scale is None:
                return ops.geglu(y, bias)
            yb = y if bias is None else y + bias
            a, b = torch.chunk(yb, 2, dim=-1)
            if cfg.activation_func_clamp_value is not None:
                a = a.clamp(max=cfg.activation_func_clamp_value)
                b = b.clamp(min=-cfg.activation_func_clamp_value, max=cfg.activation_func_clamp_value)
            out = self.activation_func(a) * (b + cfg.glu_linear_offset)
        else:
            if self.activation_func is F.gelu and cfg.bias_activation_fusion:
                out = ops.bias_gelu(y, bias)
            else:
                out = self.activation_func(y if bias is None else y + bias)
        if per_token_scale is not None:
            out = (out * per_token_scale.to(out.dtype))
        return out

    def forward(self, hidden_states, per_token_scale=None):
        inter, bias = self.linear_fc1(hidden_states)
        if "mlp_act" in (self.config.recompute_modules or []) and self.training:
            from ..tensor_parallel.random import CheckpointWithoutOutput

            ck = CheckpointWithoutOutput()
            act = ck.checkpoint(lambda y: self._activation(y, bias, per_token_scale), inter)
            out, out_bias = self.linear_fc2(act)
            ck.discard_output_and_register_recompute(out)
            return out, out_bias
        act = self._activation(inter, bias, per_token_scale)
        return self.linear_fc2(act)

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: Optional[dict] = None) -> ShardedStateDict:
        out = {}
        for name, module in self._modules.items():
            sub = sharded_state_dict_default(module, f"{prefix}{name}.", sharded_offsets, metadata)
            if self.config.gated_linear_unit and name == "linear_fc1":
                for k, v in sub.items():
                    if k in (f"{prefix}{name}.weight", f"{prefix}{name}.bias"):
                        sub[k] = apply_swiglu_sharded_factory(v, sharded_offsets)
            out.update(sub)
        return out


def apply_swiglu_sharded_factory(original_sh_ten: ShardedTensor, sharded_offsets, singleton_local_shards: bool = False):
    """The local fc1 tensor is [gate_shard ; up_shard].  Globally the layout is
    [all gate shards ; all up shards], i.e. axis 0 is fragmented 2*tp ways and this
    rank owns fragments ``tp_rank`` and ``tp + tp_rank`` (reference ``mlp.py:427-563``)."""
    swiglu_axis = 0
    prepend = len(sharded_offsets)
    frag = original_sh_ten.axis_fragmentations[prepend + swiglu_axis]
    rank_off = original_sh_ten.local_chunk_offset_in_global()[prepend + swiglu_axis]
    base_offsets = tuple(sharded_offsets)

    def build(key: str, t: torch.Tensor, replica_id: ReplicaId, flattened_range):
        w, v = torch.chunk(t, 2, dim=swiglu_axis)
        mk = lambda data, off: ShardedTensor.from_rank_offsets(  # noqa: E731
            key, data, *base_offsets, (prepend + swiglu_axis, off, frag * 2), replica_id=replica_id, prepend_axis_num=prepend
        )
        return [mk(w, rank_off), mk(v, frag + rank_off)]

    def merge(sub_state_dict):
        with torch.no_grad():
            return torch.cat(sub_state_dict)

    return ShardedTensorFactory(original_sh_ten.key, original_sh_ten.data, build, merge, original_sh_ten.replica_id)
