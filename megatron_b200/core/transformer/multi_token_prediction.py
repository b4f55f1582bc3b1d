"""Multi-token prediction (DeepSeek-V3 §2.2) — reference ``transformer/multi_token_prediction.py`` (2,170 LoC).

Depth ``k`` (1-based) predicts token ``t+1+k`` from the previous depth's hidden state at ``t`` and the embedding
of token ``t+k``::

    h_k = TransformerLayer_k( W_k · [ RMSNorm(h_{k-1}) ; RMSNorm(Emb(tok_{t+k})) ] )
    loss_k = CE( OutHead(final_norm(h_k)), labels shifted by k )

The embedding and the output head are SHARED with the main model.  The MTP losses are averaged over depths,
scaled by ``mtp_loss_scaling_factor`` and attached to the main hidden state through ``MTPLossAutoScaler`` so
that ``loss.backward()`` of the main loss also back-propagates them (no change to the training loop).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch

from ..tensor_parallel.mappings import gather_from_tensor_model_parallel_region, scatter_to_sequence_parallel_region
from ..utils import get_pg_size, get_tensor_model_parallel_group_if_none
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module
from .transformer_config import TransformerConfig


def roll_tensor(t: torch.Tensor, shifts: int = -1, dims: int = -1):
    """Shift left along ``dims`` and zero the vacated tail; returns (rolled, sum(rolled))."""
    r = torch.roll(t, shifts=shifts, dims=dims)
    idx = [slice(None)] * r.dim()
    idx[dims] = slice(shifts, None) if shifts < 0 else slice(0, shifts)
    r[tuple(idx)] = 0
    return r, r.sum()


class MTPLossAutoScaler(torch.autograd.Function):
    """Identity on ``output`` whose backward injects ``d(mtp_loss) = main_loss_backward_scale``."""

    main_loss_backward_scale: torch.Tensor = torch.tensor(1.0)

    @staticmethod
    def forward(ctx, output: torch.Tensor, mtp_loss: torch.Tensor):
        ctx.save_for_backward(mtp_loss)
        return output

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        (mtp_loss,) = ctx.saved_tensors
        scale = MTPLossAutoScaler.main_loss_backward_scale
        return grad_output, torch.ones_like(mtp_loss) * scale.to(mtp_loss.device)

    @staticmethod
    def set_loss_scale(scale: torch.Tensor):
        """Called by the schedule with ``loss_scale / num_microbatches`` (reference ``schedules.py`` MTP hook)."""
        MTPLossAutoScaler.main_loss_backward_scale = scale


class MTPLossLoggingHelper:
    """Accumulates per-depth MTP losses for the training log (reference ``:MTPLossLoggingHelper``)."""

    tracker: dict = {}

    @staticmethod
    def save_loss_to_tracker(loss: torch.Tensor, layer_number: int, num_layers: int):
        t = MTPLossLoggingHelper.tracker
        if "values" not in t:
            t["values"] = torch.zeros(num_layers, device=loss.device)
        t["values"][layer_number] += loss.detach()

    @staticmethod
    def pop(total_loss_dict: Optional[dict] = None, scale: float = 1.0) -> dict:
        t = MTPLossLoggingHelper.tracker
        out = {}
        if "values" in t:
            for i, v in enumerate(t["values"] * scale):
                out[f"mtp_{i + 1} loss"] = v
            t["values"].zero_()
        if total_loss_dict is not None:
            for k, v in out.items():
                total_loss_dict[k] = total_loss_dict.get(k, 0.0) + v
        return out


@dataclass
class MultiTokenPredictionLayerSubmodules:
    enorm: Union[ModuleSpec, type] = None
    hnorm: Union[ModuleSpec, type] = None
    eh_proj: Union[ModuleSpec, type] = None
    transformer_layer: Union[ModuleSpec, type] = None
    layer_norm: Union[ModuleSpec, type] = None


@dataclass
class MultiTokenPredictionBlockSubmodules:
    layer_specs: list = None


class MultiTokenPredictionLayer(MegatronModule):
    def __init__(self, config: TransformerConfig, submodules: MultiTokenPredictionLayerSubmodules, layer_number: int = 1, vp_stage=None):
        super().__init__(config)
        self.layer_number = layer_number
        self.tp_group = get_tensor_model_parallel_group_if_none(None)
        self.sequence_parallel = config.sequence_parallel and get_pg_size(self.tp_group) > 1
        h = config.hidden_size
        self.enorm = build_module(submodules.enorm, config=config, hidden_size=h, eps=config.layernorm_epsilon)
        self.hnorm = build_module(submodules.hnorm, config=config, hidden_size=h, eps=config.layernorm_epsilon)
        # [h_prev ; emb] (2h) -> h, column-parallel; its output is re-gathered below so the layer sees full hidden
        self.eh_proj = build_module(submodules.eh_proj, 2 * h, h, config=config, init_method=config.init_method, gather_output=False,
                                    bias=False, skip_bias_add=False, is_expert=False)
        self.transformer_layer = build_module(submodules.transformer_layer, config=config, layer_number=config.num_layers + layer_number)
        self.final_layernorm = build_module(submodules.layer_norm, config=config, hidden_size=h, eps=config.layernorm_epsilon)

    def forward(self, decoder_input, hidden_states, attention_mask, rotary_pos_emb=None, inference_context=None, packed_seq_params=None, **kw):
        x = torch.cat((self.enorm(decoder_input), self.hnorm(hidden_states)), dim=-1)
        x, _ = self.eh_proj(x)  # [s, b, h/tp]  (SP: the all-gather over s happened inside)
        x = gather_from_tensor_model_parallel_region(x, group=self.tp_group)
        if self.sequence_parallel:
            x = scatter_to_sequence_parallel_region(x, group=self.tp_group)
        out = self.transformer_layer(hidden_states=x, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb,
                                     inference_context=inference_context, packed_seq_params=packed_seq_params)
        x = out[0] if isinstance(out, tuple) else out
        return self.final_layernorm(x)


class MultiTokenPredictionBlock(MegatronModule):
    def __init__(self, config: TransformerConfig, spec: Union[MultiTokenPredictionBlockSubmodules, ModuleSpec], vp_stage=None):
        super().__init__(config)
        subs = spec.submodules if isinstance(spec, ModuleSpec) else spec
        self.layers = torch.nn.ModuleList(
            [build_module(ls, config=config, layer_number=i + 1, vp_stage=vp_stage) for i, ls in enumerate(subs.layer_specs)]
        )
        self.mtp_loss_scaling_factor = config.mtp_loss_scaling_factor

    def forward(self, input_ids, position_ids, hidden_states, attention_mask, labels=None, loss_mask=None, rotary_pos_emb=None,
                embedding=None, output_layer=None, output_weight=None, compute_language_model_loss=None, inference_context=None,
                packed_seq_params=None, **kw):
        """Returns the main hidden state (with the scaled MTP losses attached for backward when training)."""
        if labels is None:
            return hidden_states  # MTP heads are a training-time auxiliary (speculative decoding uses them explicitly)
        if loss_mask is None:
            loss_mask = torch.ones_like(labels, dtype=torch.float32)
        main_hidden = hidden_states
        h = hidden_states
        n = len(self.layers)
        ids, lbl, msk = input_ids, labels, loss_mask
        for i, layer in enumerate(self.layers):
            ids, _ = roll_tensor(ids, -1, -1)
            lbl, _ = roll_tensor(lbl, -1, -1)
            msk, n_tok = roll_tensor(msk, -1, -1)
            emb = embedding(input_ids=ids, position_ids=position_ids)
            h = layer(decoder_input=emb, hidden_states=h, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb,
                      inference_context=inference_context, packed_seq_params=packed_seq_params)
            logits, _ = output_layer(h, weight=output_weight)
            loss = compute_language_model_loss(lbl, logits)  # [b, s]
            loss = (loss * msk).sum() / n_tok.clamp(min=1)
            if self.training:
                MTPLossLoggingHelper.save_loss_to_tracker(loss, i, n)
            main_hidden = MTPLossAutoScaler.apply(main_hidden, self.mtp_loss_scaling_factor * loss / n)
        return main_hidden


def get_mtp_layer_spec(transformer_layer_spec: ModuleSpec, use_transformer_engine: bool = False) -> ModuleSpec:
    from ..tensor_parallel.layers import ColumnParallelLinear
    from .torch_norm import FusedNorm

    return ModuleSpec(
        module=MultiTokenPredictionLayer,
        submodules=MultiTokenPredictionLayerSubmodules(enorm=FusedNorm, hnorm=FusedNorm, eh_proj=ColumnParallelLinear,
                                                       transformer_layer=transformer_layer_spec, layer_norm=FusedNorm),
    )


def _layout_of(config):
    lay = getattr(config, "pipeline_model_parallel_layout", None)
    if isinstance(lay, str):
        from .pipeline_parallel_layer_layout import PipelineParallelLayerLayout

        lay = PipelineParallelLayerLayout.from_str(lay, config.pipeline_model_parallel_size)
    return lay


def mtp_on_this_rank(config: TransformerConfig, ignore_virtual: bool = True, vp_stage=None, pp_rank=None) -> bool:
    """Reference ``multi_token_prediction.py:788``: with an explicit pipeline layout the ``m`` symbols say where the MTP layers live (any stage
    from the one holding the last decoder layer onwards); without one they sit on the last stage."""
    from .. import parallel_state as ps
    from ..enums import LayerType

    if not config.mtp_num_layers:
        return False
    lay = _layout_of(config)
    if lay is None:
        return ps.is_pipeline_last_stage(ignore_virtual=ignore_virtual, vp_stage=vp_stage) if ps.model_parallel_is_initialized() else True
    r = pp_rank if pp_rank is not None else (ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0)
    stages = lay.layout[r]
    if not ignore_virtual and len(stages) > 1:
        assert vp_stage is not None, "vp_stage must be passed when virtual pipeline parallelism is on"
        return stages[vp_stage].count(LayerType.mtp) > 0
    return any(st.count(LayerType.mtp) > 0 for st in stages)


def get_mtp_num_layers_to_build(config: TransformerConfig, vp_stage=None, pp_rank=None) -> int:
    from .. import parallel_state as ps
    from ..enums import LayerType

    lay = _layout_of(config)
    if lay is not None and config.mtp_num_layers:
        r = pp_rank if pp_rank is not None else (ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0)
        return lay.get_num_layers_to_build(LayerType.mtp, vp_stage, r)
    last = ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=vp_stage) if ps.model_parallel_is_initialized() else True
    return (config.mtp_num_layers or 0) if last else 0


def get_mtp_layer_offset(config: TransformerConfig, vp_stage=None, pp_rank=None) -> int:
    """Depth index of the first MTP layer built on this (pp rank, vp stage) — MTP layers may be spread over several stages."""
    from .. import parallel_state as ps
    from ..enums import LayerType

    lay = _layout_of(config)
    if lay is None:
        return 0
    r = pp_rank if pp_rank is not None else (ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0)
    return lay.get_layer_offset(LayerType.mtp, vp_stage, r)


def get_mtp_block_spec(config: TransformerConfig, transformer_layer_spec: ModuleSpec, use_transformer_engine: bool = False, vp_stage=None):
    n = get_mtp_num_layers_to_build(config, vp_stage)
    if n == 0:
        return None
    layer = get_mtp_layer_spec(transformer_layer_spec, use_transformer_engine)
    return ModuleSpec(module=MultiTokenPredictionBlock, submodules=MultiTokenPredictionBlockSubmodules(layer_specs=[layer] * n))
