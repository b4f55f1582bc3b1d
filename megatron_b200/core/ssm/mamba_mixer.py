"""Mamba-2 mixer (reference ``ssm/mamba_mixer.py:143``).

    in_proj (column-parallel) → [z | x | B | C | dt]  → causal depthwise conv on [x|B|C] → SSD scan → gated RMSNorm(y·silu(z)) → out_proj (row-parallel)

Heads (and the B/C groups) are sharded over TP; the conv and the scan are purely local.  Decoding keeps a conv window and the
SSM state per layer in ``inference_context.key_value_memory_dict``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import torch

from ..utils import divide, get_pg_size, get_tensor_model_parallel_group_if_none
from ..transformer.module import MegatronModule
from ..transformer.spec_utils import ModuleSpec, build_module
from ..transformer.transformer_config import TransformerConfig
from .ssd import causal_conv1d, causal_conv1d_update, ssd_chunk_scan, ssd_step


@dataclass
class MambaMixerSubmodules:
    in_proj: Union[ModuleSpec, type] = None
    out_proj: Union[ModuleSpec, type] = None


class MambaMixer(MegatronModule):
    def __init__(self, config: TransformerConfig, submodules: MambaMixerSubmodules, d_model: int, d_conv: int = 4, conv_init=None,
                 expand: int = 2, A_init_range=(1, 16), D_has_hdim: bool = False, rmsnorm: bool = True, norm_before_gate: bool = False,
                 dt_min: float = 0.001, dt_max: float = 0.1, dt_init_floor: float = 1e-4, bias: bool = False, conv_bias: bool = True,
                 chunk_size: int = 128, layer_number: Optional[int] = None, pg_collection=None, **_):
        super().__init__(config)
        self.layer_number = layer_number
        self.tp_group = pg_collection.tp if pg_collection is not None and getattr(pg_collection, "tp", None) is not None else get_tensor_model_parallel_group_if_none(None)
        ws = get_pg_size(self.tp_group)
        self.d_model, self.d_conv, self.chunk_size = d_model, d_conv, chunk_size
        self.d_state = getattr(config, "mamba_state_dim", 128)
        self.headdim = getattr(config, "mamba_head_dim", 64)
        self.ngroups = getattr(config, "mamba_num_groups", 8)
        nheads_cfg = getattr(config, "mamba_num_heads", None)
        self.d_inner = nheads_cfg * self.headdim if nheads_cfg else expand * d_model
        self.nheads = divide(self.d_inner, self.headdim)
        self.rmsnorm, self.norm_before_gate, self.D_has_hdim = rmsnorm, norm_before_gate, D_has_hdim
        self.nheads_local, self.ngroups_local, self.d_inner_local = divide(self.nheads, ws), divide(self.ngroups, ws), divide(self.d_inner, ws)
        assert self.nheads_local % self.ngroups_local == 0
        proj_dim = 2 * self.d_inner + 2 * self.ngroups * self.d_state + self.nheads
        self.in_proj = build_module(submodules.in_proj, d_model, proj_dim, config=config, init_method=config.init_method, gather_output=False,
                                    bias=bias, skip_bias_add=False, is_expert=False, tp_group=self.tp_group)
        conv_dim = self.d_inner_local + 2 * self.ngroups_local * self.d_state
        dev = self.in_proj.weight.device
        dt_ = config.params_dtype
        self.conv1d_weight = torch.nn.Parameter(torch.empty(conv_dim, d_conv, device=dev, dtype=dt_))
        self.conv1d_bias = torch.nn.Parameter(torch.zeros(conv_dim, device=dev, dtype=dt_)) if conv_bias else None
        with torch.no_grad():
            bound = conv_init if conv_init is not None else 1.0 / math.sqrt(d_conv)
            self.conv1d_weight.uniform_(-bound, bound)
            dt = torch.exp(torch.rand(self.nheads_local, device=dev) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
            inv_dt = dt + torch.log(-torch.expm1(-dt))
            A = torch.empty(self.nheads_local, device=dev).uniform_(*A_init_range)
        self.dt_bias = torch.nn.Parameter(inv_dt.float())
        self.A_log = torch.nn.Parameter(torch.log(A).float())
        self.D = torch.nn.Parameter(torch.ones(self.d_inner_local if D_has_hdim else self.nheads_local, device=dev, dtype=torch.float32))
        for prm in (self.conv1d_weight, self.conv1d_bias, self.dt_bias, self.A_log, self.D):
            if prm is not None:
                setattr(prm, "tensor_model_parallel", True)
                setattr(prm, "partition_dim", 0)
                setattr(prm, "partition_stride", 1)
        self.dt_bias._no_weight_decay = self.A_log._no_weight_decay = self.D._no_weight_decay = True
        if rmsnorm:
            self.norm_weight = torch.nn.Parameter(torch.ones(self.d_inner_local, device=dev, dtype=dt_))
            setattr(self.norm_weight, "tensor_model_parallel", True)
            setattr(self.norm_weight, "partition_dim", 0)
        self.out_proj = build_module(submodules.out_proj, self.d_inner, d_model, config=config, init_method=config.output_layer_init_method,
                                     bias=bias, input_is_parallel=True, skip_bias_add=True, is_expert=False, tp_group=self.tp_group)

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def _gated_norm(self, y, z):
        """RMSNorm over the per-group slice of d_inner with SiLU gate (reference ``RMSNormGated`` with group_size = d_inner/ngroups)."""
        if not self.rmsnorm:
            return y * torch.nn.functional.silu(z)
        gs = self.d_inner_local // self.ngroups_local
        if not self.norm_before_gate:
            y = y * torch.nn.functional.silu(z)
        yf = y.float().view(*y.shape[:-1], self.ngroups_local, gs)
        yf = yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + self.config.layernorm_epsilon)
        out = (yf.view(*y.shape) * self.norm_weight.float()).to(y.dtype)
        if self.norm_before_gate:
            out = out * torch.nn.functional.silu(z)
        return out

    def _split(self, zxbcdt):
        di, gn, nh = self.d_inner_local, self.ngroups_local * self.d_state, self.nheads_local
        return torch.split(zxbcdt, [di, di + 2 * gn, nh], dim=-1)

    def forward(self, hidden_states, inference_context=None, *, inference_params=None, **_):
        """hidden_states [s(/tp under SP), b, h] → ([s(/tp), b, h], bias)."""
        inference_context = inference_context or inference_params
        zxbcdt, _ = self.in_proj(hidden_states)                      # [l, b, proj_local]
        l, b = zxbcdt.shape[:2]
        A = -torch.exp(self.A_log.float())
        z, xBC, dt = self._split(zxbcdt)
        di, gn = self.d_inner_local, self.ngroups_local * self.d_state
        D = self.D.view(self.nheads_local, -1) if self.D_has_hdim else self.D
        if inference_context is not None and inference_context.sequence_len_offset > 0 and l == 1:
            conv_state, ssm_state = inference_context.key_value_memory_dict[("mamba", self.layer_number)]
            xBC1 = causal_conv1d_update(xBC[0], conv_state, self.conv1d_weight, self.conv1d_bias)
            x, B, C = torch.split(xBC1, [di, gn, gn], dim=-1)
            dt1 = torch.nn.functional.softplus(dt[0].float() + self.dt_bias)
            y, new_state = ssd_step(x.view(b, self.nheads_local, self.headdim), dt1, A, B.view(b, self.ngroups_local, self.d_state),
                                    C.view(b, self.ngroups_local, self.d_state), ssm_state, D)
            ssm_state.copy_(new_state)
            y = y.reshape(1, b, di)
        elif self._cp_size() > 1:
            y = self._forward_context_parallel(z, xBC, dt, A, D)
            z = None
        else:
            xc = xBC.permute(1, 2, 0)                                # [b, conv_dim, l]
            want_state = inference_context is not None
            conv_out = causal_conv1d(xc, self.conv1d_weight, self.conv1d_bias, "silu", return_final_state=want_state)
            if want_state:
                conv_out, conv_state = conv_out
            xBCc = conv_out.permute(0, 2, 1)                         # [b, l, conv_dim]
            x, B, C = torch.split(xBCc, [di, gn, gn], dim=-1)
            dtp = torch.nn.functional.softplus(dt.float().permute(1, 0, 2) + self.dt_bias)   # [b, l, nh]
            res = ssd_chunk_scan(x.reshape(b, l, self.nheads_local, self.headdim), dtp, A, B.reshape(b, l, self.ngroups_local, self.d_state),
                                 C.reshape(b, l, self.ngroups_local, self.d_state), self.chunk_size, D, return_final_states=want_state)
            if want_state:
                res, ssm_state = res
                inference_context.key_value_memory_dict[("mamba", self.layer_number)] = (conv_state.clone(), ssm_state.clone())
            y = res.reshape(b, l, di).permute(1, 0, 2)               # [l, b, d_inner_local]
        if z is not None:
            y = self._gated_norm(y, z)
        return self.out_proj(y)

    # ---- context parallel ---------------------------------------------------------------------------------------------
    def _cp_size(self) -> int:
        from .. import parallel_state as ps

        return ps.get_context_parallel_world_size() if ps.is_initialized() else 1

    def _forward_context_parallel(self, z, xBC, dt, A, D):
        """Sequence-sharded activations are re-sharded to head blocks (one all-to-all per tensor), the conv + scan run over the FULL sequence for ``1/cp`` of the
        heads / groups with the matching parameter slices, and the result goes back to the sequence sharding.  Parameters stay replicated over CP; every
        rank only produces gradients for its slice and the DP×CP gradient reduction adds them up."""
        from .. import parallel_state as ps
        from ...parallel.context_parallel import channel_to_seq, seq_to_channel

        group, cp, r = ps.get_context_parallel_group(), ps.get_context_parallel_world_size(), ps.get_context_parallel_rank()
        di, ng, nh, ds, hd = self.d_inner_local, self.ngroups_local, self.nheads_local, self.d_state, self.headdim
        assert nh % cp == 0 and ng % cp == 0, "Mamba context parallelism needs heads and groups divisible by the CP size"
        x, B, C = torch.split(xBC, [di, ng * ds, ng * ds], dim=-1)
        zc, xc, Bc, Cc, dtc = (seq_to_channel(t.contiguous(), group, cp) for t in (z, x, B, C, dt))
        dil, ngl, nhl = di // cp, ng // cp, nh // cp
        xs, bs_, cs_ = slice(r * dil, (r + 1) * dil), slice(di + r * ngl * ds, di + (r + 1) * ngl * ds), slice(di + ng * ds + r * ngl * ds, di + ng * ds + (r + 1) * ngl * ds)
        w = torch.cat([self.conv1d_weight[xs], self.conv1d_weight[bs_], self.conv1d_weight[cs_]], dim=0)
        bias = None if self.conv1d_bias is None else torch.cat([self.conv1d_bias[xs], self.conv1d_bias[bs_], self.conv1d_bias[cs_]], dim=0)
        l, b = xc.shape[:2]
        conv_in = torch.cat([xc, Bc, Cc], dim=-1).permute(1, 2, 0)                        # [b, conv_dim/cp, l]
        conv_out = causal_conv1d(conv_in, w, bias, "silu").permute(0, 2, 1)              # [b, l, conv_dim/cp]
        xl, Bl, Cl = torch.split(conv_out, [dil, ngl * ds, ngl * ds], dim=-1)
        hs = slice(r * nhl, (r + 1) * nhl)
        dtp = torch.nn.functional.softplus(dtc.float().permute(1, 0, 2) + self.dt_bias[hs])
        Dl = (D.view(nh, -1)[hs] if self.D_has_hdim else D[hs])
        y = ssd_chunk_scan(xl.reshape(b, l, nhl, hd), dtp, A[hs], Bl.reshape(b, l, ngl, ds), Cl.reshape(b, l, ngl, ds), self.chunk_size, Dl)
        y = y.reshape(b, l, dil).permute(1, 0, 2)                                        # [l, b, d_inner/cp]
        # gated norm on the local group block
        if not self.rmsnorm:
            y = y * torch.nn.functional.silu(zc)
        else:
            gs = di // ng
            if not self.norm_before_gate:
                y = y * torch.nn.functional.silu(zc)
            yf = y.float().view(l, b, ngl, gs)
            yf = yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + self.config.layernorm_epsilon)
            y = (yf.view(l, b, dil) * self.norm_weight[xs].float()).to(y.dtype)
            if self.norm_before_gate:
                y = y * torch.nn.functional.silu(zc)
        return channel_to_seq(y.contiguous(), group, cp)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..transformer.utils import make_sharded_tensors_for_checkpoint, sharded_state_dict_default

        sd = {}
        own = {k: v for k, v in self.state_dict(prefix="", keep_vars=True).items() if "." not in k}
        sd.update(make_sharded_tensors_for_checkpoint(own, prefix, {k: 0 for k in own}, sharded_offsets, tp_group=self.tp_group))
        for name, mod in (("in_proj", self.in_proj), ("out_proj", self.out_proj)):
            sd.update(sharded_state_dict_default(mod, f"{prefix}{name}.", sharded_offsets, metadata))
        return sd
