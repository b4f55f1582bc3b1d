"""One transformer layer (reference ``transformer/transformer_layer.py:313``).

Pre-LN residual block: ``x + BDA(attn(LN(x)))`` then ``x + BDA(mlp(LN(x)))``.
Under sequence parallelism the layer input/output stay sharded ``[s/tp, b, h]``;
the all-gather / reduce-scatter live inside the linear pair ops.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Union

import torch

from ... import ops
from .. import parallel_state as ps
from ..utils import make_viewless_tensor, nvtx_range_pop, nvtx_range_push
from .identity_op import IdentityFuncOp, IdentityOp
from .module import GraphableMegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import TransformerConfig
from .utils import sharded_state_dict_default


def get_transformer_layer_offset(config: TransformerConfig, vp_stage: Optional[int] = None, pp_rank: Optional[int] = None) -> int:
    """Global index of the first layer owned by this (pp_rank, vp_stage).

    Interleaved layout: stage ``p`` chunk ``v`` holds layers
    ``[v * (L/vp) + p * (L/(pp*vp)), ...)`` (reference :65-311, uniform split incl.
    uneven first/last stage and embedding/loss accounting).
    """
    pp = config.pipeline_model_parallel_size
    if pp <= 1:
        return 0
    pp_rank = ps.get_pipeline_model_parallel_rank() if pp_rank is None else pp_rank
    vp = config.virtual_pipeline_model_parallel_size
    first, last = config.num_layers_in_first_pipeline_stage, config.num_layers_in_last_pipeline_stage
    n = config.num_layers
    if first is not None or last is not None:
        mid_stages = pp - (first is not None) - (last is not None)
        n_mid = n - (first or 0) - (last or 0)
        per_mid = n_mid // mid_stages if mid_stages > 0 else 0
        vpn = vp or 1
        v = vp_stage or 0
        f_per = (first // vpn) if first is not None else per_mid // vpn
        l_per = (last // vpn) if last is not None else per_mid // vpn
        m_per = per_mid // vpn
        per_chunk_total = f_per + l_per + m_per * mid_stages if (first is not None and last is not None) else None
        # layers in one virtual chunk across all stages
        chunk_total = (f_per if first is not None else m_per) + (l_per if last is not None else m_per) + m_per * (pp - 2)
        off = v * chunk_total
        if pp_rank == 0:
            return off
        off += f_per if first is not None else m_per
        off += m_per * (pp_rank - 1)
        return off
    total = n
    if config.account_for_embedding_in_pipeline_split:
        total += 1
    if config.account_for_loss_in_pipeline_split:
        total += 1
    per_stage = total // pp
    if vp is not None:
        v = vp_stage if vp_stage is not None else (ps.get_virtual_pipeline_model_parallel_rank() or 0)
        per_chunk = per_stage // vp
        off = v * (total // vp) + pp_rank * per_chunk
    else:
        off = pp_rank * per_stage
    if config.account_for_embedding_in_pipeline_split and not (pp_rank == 0 and (vp is None or (vp_stage or 0) == 0)):
        off -= 1
    return off


@dataclass
class TransformerLayerSubmodules:
    input_layernorm: Union[ModuleSpec, type] = IdentityOp
    self_attention: Union[ModuleSpec, type] = IdentityOp
    self_attn_bda: Union[ModuleSpec, type] = IdentityFuncOp
    pre_cross_attn_layernorm: Union[ModuleSpec, type] = IdentityOp
    cross_attention: Union[ModuleSpec, type] = IdentityOp
    cross_attn_bda: Union[ModuleSpec, type] = IdentityFuncOp
    pre_mlp_layernorm: Union[ModuleSpec, type] = IdentityOp
    mlp: Union[ModuleSpec, type] = IdentityOp
    mlp_bda: Union[ModuleSpec, type] = IdentityFuncOp
    sharded_state_dict_keys_map: Dict[str, str] = field(default_factory=dict)


class BaseTransformerLayer:
    """Marker base so custom layers can be recognised by the block."""


class TransformerLayer(GraphableMegatronModule, BaseTransformerLayer):
    def __init__(self, config: TransformerConfig, submodules: TransformerLayerSubmodules, layer_number: int = 1,
                 hidden_dropout: Optional[float] = None, pg_collection=None, vp_stage: Optional[int] = None, **kwargs):
        super().__init__(config, vp_stage=vp_stage)
        self.submodules_config = submodules
        self.layer_number = layer_number + get_transformer_layer_offset(config, vp_stage)
        self.hidden_dropout = config.hidden_dropout if hidden_dropout is None else hidden_dropout
        self.pg_collection = pg_collection
        norm_kw = dict(config=config, hidden_size=config.hidden_size, eps=config.layernorm_epsilon)
        self.input_layernorm = build_module(submodules.input_layernorm, **norm_kw)
        attn_kw = {"pg_collection": pg_collection} if pg_collection is not None else {}
        self.self_attention = build_module(submodules.self_attention, config=config, layer_number=self.layer_number, **attn_kw)
        self.self_attn_bda = build_module(submodules.self_attn_bda)
        self.pre_cross_attn_layernorm = build_module(submodules.pre_cross_attn_layernorm, **norm_kw)
        self.cross_attention = build_module(submodules.cross_attention, config=config, layer_number=self.layer_number, **attn_kw)
        self.cross_attn_bda = build_module(submodules.cross_attn_bda, config=config) if submodules.cross_attn_bda is not IdentityFuncOp else build_module(submodules.cross_attn_bda)
        self.pre_mlp_layernorm = build_module(submodules.pre_mlp_layernorm, **norm_kw)
        mlp_kw = {}
        from .moe.moe_layer import MoELayer  # local import: moe depends on this package

        mlp_mod = submodules.mlp.module if isinstance(submodules.mlp, ModuleSpec) else submodules.mlp
        self.is_moe_layer = isinstance(mlp_mod, type) and issubclass(mlp_mod, MoELayer)
        if self.is_moe_layer:
            mlp_kw = dict(layer_number=self.layer_number, pg_collection=pg_collection)
        self.mlp = build_module(submodules.mlp, config=config, **mlp_kw)
        if hasattr(self.mlp, "set_layer_number"):
            self.mlp.set_layer_number(self.layer_number)
        self.mlp_bda = build_module(submodules.mlp_bda)
        self.mhc = bool(getattr(config, "enable_mhc_connections", False))
        if self.mhc:
            from .hyper_connection import HyperConnectionModule

            self.self_attention_hyper_connection = HyperConnectionModule(config, self.layer_number)
            self.mlp_hyper_connection = HyperConnectionModule(config, self.layer_number)
        rm = set(config.recompute_modules or []) if config.recompute_granularity == "selective" else set()
        self.recompute_input_layernorm = "layernorm" in rm and not isinstance(self.input_layernorm, IdentityOp)
        self.recompute_pre_mlp_layernorm = "layernorm" in rm and not isinstance(self.pre_mlp_layernorm, IdentityOp)
        self.recompute_mlp = "mlp" in rm or ("moe" in rm and self.is_moe_layer)

    # -- helpers -----------------------------------------------------------------
    def _bda(self, bda, out_with_bias, residual):
        out, bias = out_with_bias
        fn = bda(self.training, self.config.bias_dropout_fusion) if callable(bda) else None
        if fn is None or isinstance(bda, IdentityFuncOp):
            return out if residual is None else out
        return fn((out, bias), residual, self.hidden_dropout)

    def _can_fuse_residual_norm(self, attn_out) -> bool:
        from .torch_norm import FusedNorm

        n = self.pre_mlp_layernorm
        return (self.config.fused_residual_rmsnorm and isinstance(n, FusedNorm) and n.normalization == "RMSNorm" and attn_out[1] is None
                and (self.hidden_dropout == 0.0 or not self.training) and not self.recompute_pre_mlp_layernorm and isinstance(self.cross_attention, IdentityOp))

    def _norm_maybe_recompute(self, norm, x, flag):
        if flag and self.training:
            from ..tensor_parallel.random import CheckpointWithoutOutput

            ck = CheckpointWithoutOutput()
            return ck.checkpoint(norm, x), ck
        return norm(x), None

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None, rotary_pos_emb=None,
                rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None, inference_context=None, packed_seq_params=None,
                sequence_len_offset=None, *, inference_params=None, **kwargs):
        inference_context = inference_context or inference_params
        hidden_states, context = self._forward_attention(
            hidden_states, attention_mask, context, context_mask, rotary_pos_emb, attention_bias, inference_context, packed_seq_params
        )
        output = self._forward_mlp(hidden_states)
        return output, context

    def _forward_attention(self, hidden_states, attention_mask, context, context_mask, rotary_pos_emb, attention_bias, inference_context, packed_seq_params):
        nvtx_range_push("attn")
        if self.mhc:
            # n-wide residual stream: the layer reads Σ h_pre_i x_i and writes H_res x + H_postᵀ F(·) (hyper_connection.HyperConnectionModule)
            hc = self.self_attention_hyper_connection
            residual = hidden_states
            x, h_res, h_post = hc(hidden_states)
            attn_out = self.self_attention(self.input_layernorm(x), attention_mask=attention_mask, inference_context=inference_context, rotary_pos_emb=rotary_pos_emb,
                                           attention_bias=attention_bias, packed_seq_params=packed_seq_params)
            hidden_states = hc.fused_h_res_h_post_bda(h_res, residual, h_post, attn_out, self.hidden_dropout, self.training)
            nvtx_range_pop("attn")
            return hidden_states, context
        residual = hidden_states
        normed, ck = self._norm_maybe_recompute(self.input_layernorm, hidden_states, self.recompute_input_layernorm)
        attn_out = self.self_attention(
            normed, attention_mask=attention_mask, inference_context=inference_context, rotary_pos_emb=rotary_pos_emb,
            attention_bias=attention_bias, packed_seq_params=packed_seq_params,
        )
        if ck is not None:
            ck.discard_output_and_register_recompute(attn_out[0])
        if self._can_fuse_residual_norm(attn_out):
            # h = residual + attn_out and pre_mlp_layernorm(h) in ONE pass; _forward_mlp picks the normed tensor up instead of re-reading h
            from ... import ops

            n = self.pre_mlp_layernorm
            self._prenormed, hidden_states = ops.add_rms_norm(attn_out[0], residual, n.weight, n.eps, n.zero_centered_gamma)
        else:
            hidden_states = self.self_attn_bda(self.training, self.config.bias_dropout_fusion)(attn_out, residual, self.hidden_dropout)
        if not isinstance(self.cross_attention, IdentityOp):
            residual = hidden_states
            normed = self.pre_cross_attn_layernorm(hidden_states)
            xo = self.cross_attention(normed, attention_mask=context_mask, key_value_states=context, inference_context=inference_context)
            if isinstance(xo, dict) and "context" in xo:
                context = xo["context"]
            hidden_states = self.cross_attn_bda(self.training, self.config.bias_dropout_fusion)(xo, residual, self.hidden_dropout)
        nvtx_range_pop("attn")
        return hidden_states, context

    def _forward_mlp(self, hidden_states):
        nvtx_range_push("mlp")
        if self.mhc:
            hc = self.mlp_hyper_connection
            residual = hidden_states
            x, h_res, h_post = hc(hidden_states)
            mlp_out = self.mlp(self.pre_mlp_layernorm(x))
            hidden_states = hc.fused_h_res_h_post_bda(h_res, residual, h_post, mlp_out, self.hidden_dropout, self.training)
            nvtx_range_pop("mlp")
            return make_viewless_tensor(hidden_states, requires_grad=hidden_states.requires_grad, keep_graph=True)
        residual = hidden_states
        pre = getattr(self, "_prenormed", None)
        if pre is not None:
            normed, ck, self._prenormed = pre, None, None
        else:
            normed, ck = self._norm_maybe_recompute(self.pre_mlp_layernorm, hidden_states, self.recompute_pre_mlp_layernorm)
        if self.recompute_mlp and self.training:
            from ..tensor_parallel.random import checkpoint

            mlp_out = checkpoint(self.mlp, False, normed)
        elif self.config.mlp_chunks_for_training > 1 and self.training and not self.is_moe_layer:
            n = self.config.mlp_chunks_for_training
            parts = [self.mlp(c) for c in normed.chunk(n, dim=0)]
            mlp_out = (torch.cat([p[0] for p in parts], dim=0), parts[0][1])
        else:
            mlp_out = self.mlp(normed)
        if ck is not None:
            ck.discard_output_and_register_recompute(mlp_out[0])
        hidden_states = self.mlp_bda(self.training, self.config.bias_dropout_fusion)(mlp_out, residual, self.hidden_dropout)
        nvtx_range_pop("mlp")
        return make_viewless_tensor(hidden_states, requires_grad=hidden_states.requires_grad, keep_graph=True)

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: Optional[dict] = None):
        sd = {}
        for name, child in self.named_children():
            sd.update(sharded_state_dict_default(child, f"{prefix}{name}.", sharded_offsets, metadata))
        km = self.submodules_config.sharded_state_dict_keys_map
        if km:
            pm = {f"{prefix}{k}": f"{prefix}{v}" for k, v in km.items()}
            out = {}
            for k, v in sd.items():
                for old, new in pm.items():
                    if k.startswith(old):
                        # only the *checkpoint* key is remapped; the dict key stays the module path
                        if hasattr(v, "key"):
                            v.key = v.key.replace(old, new, 1)
                        break
                out[k] = v
            sd = out
        return sd


def get_bias_dropout_add(training: bool, fused: bool):
    """Returns ``f((x, bias), residual, prob)`` (reference ``fusions/fused_bias_dropout.py``)."""

    def f(x_with_bias, residual, prob):
        x, bias = x_with_bias
        return ops.bias_dropout_add(x, bias, residual, prob, training)

    return f


class HyperConnectionTransformerLayer(TransformerLayer):
    """``TransformerLayer`` with the n-wide mHC residual stream (reference ``transformer_layer.py`` keeps this as a subclass; here the base class already switches
    on ``config.enable_mhc_connections`` — the subclass only insists that the switch is on, so specs written for the reference resolve to the same behaviour)."""

    def __init__(self, config, *args, **kwargs):
        if not getattr(config, "enable_mhc_connections", False):
            raise ValueError("HyperConnectionTransformerLayer requires config.enable_mhc_connections=True")
        super().__init__(config, *args, **kwargs)
