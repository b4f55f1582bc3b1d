"""Gated DeltaNet linear attention (Qwen3-Next style; reference ``ssm/gated_delta_net.py`` + ``ssm/ops/gdp/*.py`` Triton kernels).

Per head, with state ``S ∈ R^{dv×dk}``:

    S_t = α_t · S_{t-1} (I − β_t k_t k_tᵀ) + β_t v_t k_tᵀ ,        o_t = S_t q_t

(α_t = exp(g_t) ∈ (0,1] is the gate/decay, β_t ∈ (0,1) the write strength, q/k l2-normalised).  ``gated_delta_rule_chunked`` evaluates it
chunk-wise with the WY representation — inside a chunk the product of (I − β k kᵀ) factors is expressed through a triangular solve, so
the work is dense matmuls of chunk size C; only the ``[dv, dk]`` state crosses chunk boundaries.  ``gated_delta_rule_recurrent`` is the
token-by-token definition (decode path and test oracle)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch
import torch.nn.functional as F

from ..transformer.module import MegatronModule
from ..transformer.spec_utils import ModuleSpec, build_module
from ..utils import divide, get_pg_size, get_tensor_model_parallel_group_if_none
from .ssd import causal_conv1d, causal_conv1d_update


def gated_delta_rule_recurrent(q, k, v, g, beta, initial_state=None):
    """q,k [b,l,h,dk]; v [b,l,h,dv]; g [b,l,h] (log-decay ≤ 0); beta [b,l,h] → (o [b,l,h,dv], final state [b,h,dv,dk])."""
    b, l, h, dk = q.shape
    dv = v.shape[-1]
    S = torch.zeros(b, h, dv, dk, dtype=torch.float32, device=q.device) if initial_state is None else initial_state.float()
    qf, kf, vf, gf, bf = q.float(), k.float(), v.float(), g.float(), beta.float()
    outs = []
    for t in range(l):
        kt, vt, bt = kf[:, t], vf[:, t], bf[:, t].unsqueeze(-1)
        S = S * torch.exp(gf[:, t])[..., None, None]
        pred = torch.einsum("bhvk,bhk->bhv", S, kt)                     # what the memory currently returns for k_t
        S = S + torch.einsum("bhv,bhk->bhvk", bt * (vt - pred), kt)     # delta rule: write the correction
        outs.append(torch.einsum("bhvk,bhk->bhv", S, qf[:, t]))
    return torch.stack(outs, dim=1).to(v.dtype), S


def gated_delta_rule_chunked(q, k, v, g, beta, chunk_size: int = 64, initial_state=None):
    """Chunk-parallel evaluation of the same recurrence (matmul-shaped work; exact up to fp32 round-off)."""
    b, l, h, dk = q.shape
    dv = v.shape[-1]
    pad = (-l) % chunk_size
    if pad:
        q, k, v = (F.pad(t, (0, 0, 0, 0, 0, pad)) for t in (q, k, v))
        g, beta = F.pad(g, (0, 0, 0, pad)), F.pad(beta, (0, 0, 0, pad))
    L = l + pad
    n, C = L // chunk_size, chunk_size
    f = lambda t: t.float().view(b, n, C, h, -1).permute(0, 3, 1, 2, 4)   # [b, h, n, C, d]  # noqa: E731
    qf, kf, vf = f(q), f(k), f(v)
    gf = g.float().view(b, n, C, h).permute(0, 3, 1, 2)                  # [b, h, n, C]
    bf = beta.float().view(b, n, C, h).permute(0, 3, 1, 2)
    gc = gf.cumsum(-1)                                                    # in-chunk cumulative log-decay
    strict = torch.tril(torch.ones(C, C, dtype=torch.bool, device=q.device), -1)
    incl = torch.tril(torch.ones(C, C, dtype=torch.bool, device=q.device), 0)
    # [.., i, j] = exp(gc_i − gc_j) for j <= i.  The upper triangle is masked BEFORE the exponential: there the exponent is positive and overflows for
    # strong decays, and an inf that is only zeroed afterwards turns into NaN in the backward (inf * 0)
    decay = torch.exp((gc.unsqueeze(-1) - gc.unsqueeze(-2)).masked_fill(~incl, float("-inf")))
    kb = kf * bf.unsqueeze(-1)
    # (I + A) u = β v  and  (I + A) w = β k·exp(gc), with A_ij = β_i (k_i·k_j) exp(gc_i − gc_j) for j < i   (WY representation)
    A = (torch.einsum("bhnid,bhnjd->bhnij", kb, kf) * decay).masked_fill(~strict, 0.0)
    eye = torch.eye(C, device=q.device).expand_as(A)
    Tinv = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
    u = Tinv @ (vf * bf.unsqueeze(-1))                                    # [b,h,n,C,dv]
    w = Tinv @ (kb * torch.exp(gc).unsqueeze(-1))                         # [b,h,n,C,dk]
    S = torch.zeros(b, h, dv, dk, dtype=torch.float32, device=q.device) if initial_state is None else initial_state.float()
    qk = (torch.einsum("bhnid,bhnjd->bhnij", qf, kf) * decay).masked_fill(~incl, 0.0)
    outs = []
    for c in range(n):
        v_new = u[:, :, c] - w[:, :, c] @ S.transpose(-1, -2)             # corrected values given the carried-in state
        o = (qf[:, :, c] * torch.exp(gc[:, :, c]).unsqueeze(-1)) @ S.transpose(-1, -2) + qk[:, :, c] @ v_new
        outs.append(o)
        tail = torch.exp(gc[:, :, c, -1:] - gc[:, :, c])                   # decay from each position to the chunk end
        S = S * torch.exp(gc[:, :, c, -1])[..., None, None] + torch.einsum("bhcv,bhck->bhvk", v_new * tail.unsqueeze(-1), kf[:, :, c])
    o = torch.stack(outs, dim=2).permute(0, 2, 3, 1, 4).reshape(b, L, h, dv)[:, :l]
    return o.to(v.dtype), S


@dataclass
class GatedDeltaNetSubmodules:
    in_proj: Union[ModuleSpec, type] = None
    out_proj: Union[ModuleSpec, type] = None


class GatedDeltaNet(MegatronModule):
    """Mixer: in_proj → [q | k | v | z | β | g] ; short causal conv on q,k,v ; l2norm(q,k) ; gated delta rule ; RMSNorm(o)·silu(z) ; out_proj."""

    def __init__(self, config, submodules: GatedDeltaNetSubmodules, d_model: Optional[int] = None, layer_number: Optional[int] = None, num_heads: Optional[int] = None,
                 head_k_dim: int = 128, head_v_dim: int = 128, conv_kernel: int = 4, chunk_size: int = 64, pg_collection=None, num_value_heads: Optional[int] = None, **_):
        super().__init__(config)
        d_model = d_model or config.hidden_size
        # the reference's ``--linear-*`` options (``transformer_config.linear_*``) win over the constructor defaults when the config carries them
        g = lambda n: getattr(config, n, None)  # noqa: E731
        head_k_dim, head_v_dim, conv_kernel = g("linear_key_head_dim") or head_k_dim, g("linear_value_head_dim") or head_v_dim, g("linear_conv_kernel_dim") or conv_kernel
        num_heads, num_value_heads = g("linear_num_key_heads") or num_heads, g("linear_num_value_heads") or num_value_heads
        self.layer_number, self.chunk_size, self.conv_kernel = layer_number, chunk_size, conv_kernel
        self.tp_group = pg_collection.tp if pg_collection is not None and getattr(pg_collection, "tp", None) is not None else get_tensor_model_parallel_group_if_none(None)
        ws = get_pg_size(self.tp_group)
        self.h = num_heads or config.num_attention_heads                     # key heads; TP shards them, each carries ``r`` value heads
        self.r = divide(num_value_heads, self.h) if num_value_heads else 1
        self.h_local = divide(self.h, ws)
        self.dk, self.dv = head_k_dim, head_v_dim
        proj = self.h * (2 * self.dk + self.r * (2 * self.dv + 2))
        self.in_proj = build_module(submodules.in_proj, d_model, proj, config=config, init_method=config.init_method, gather_output=False, bias=False,
                                    skip_bias_add=False, is_expert=False, tp_group=self.tp_group)
        dev = self.in_proj.weight.device
        conv_dim = self.h_local * (2 * self.dk + self.r * self.dv)
        self.conv_weight = torch.nn.Parameter(torch.empty(conv_dim, conv_kernel, device=dev, dtype=config.params_dtype).uniform_(-0.5, 0.5))
        self.A_log = torch.nn.Parameter(torch.log(torch.empty(self.h_local * self.r, device=dev).uniform_(1, 16)).float())
        self.dt_bias = torch.nn.Parameter(torch.zeros(self.h_local * self.r, device=dev, dtype=torch.float32))
        self.norm_weight = torch.nn.Parameter(torch.ones(self.dv, device=dev, dtype=config.params_dtype))
        for p in (self.conv_weight, self.A_log, self.dt_bias):
            setattr(p, "tensor_model_parallel", True)
            setattr(p, "partition_dim", 0)
        self.out_proj = build_module(submodules.out_proj, self.h * self.r * self.dv, d_model, config=config, init_method=config.output_layer_init_method, bias=False,
                                     input_is_parallel=True, skip_bias_add=True, is_expert=False, tp_group=self.tp_group)

    def forward(self, hidden_states, inference_context=None, *, inference_params=None, **_):
        inference_context = inference_context or inference_params
        x, _ = self.in_proj(hidden_states)                              # [l, b, h_local*(2dk+2dv+2)]
        l, b = x.shape[:2]
        hl, dk, dv, r = self.h_local, self.dk, self.dv, self.r
        x = x.view(l, b, hl, 2 * dk + r * (2 * dv + 2))
        qkv, z, beta_raw, g_raw = torch.split(x, [2 * dk + r * dv, r * dv, r, r], dim=-1)
        qkv = qkv.permute(1, 2, 3, 0).reshape(b, hl * (2 * dk + r * dv), l)   # [b, conv_dim, l]
        decode = inference_context is not None and inference_context.sequence_len_offset > 0 and l == 1
        key = ("gdn", self.layer_number)
        if decode:
            conv_state, S = inference_context.key_value_memory_dict[key]
            qkv = causal_conv1d_update(qkv[..., 0], conv_state, self.conv_weight).unsqueeze(-1)
        else:
            qkv = causal_conv1d(qkv, self.conv_weight, None, "silu", return_final_state=inference_context is not None)
            if inference_context is not None:
                qkv, conv_state = qkv
        qkv = qkv.view(b, hl, 2 * dk + r * dv, l).permute(0, 3, 1, 2)     # [b, l, h, ·]
        q, k, v = torch.split(qkv, [dk, dk, r * dv], dim=-1)
        q = F.normalize(q.float(), dim=-1).to(v.dtype) * (dk ** -0.5)
        k = F.normalize(k.float(), dim=-1).to(v.dtype)
        if r > 1:                                                          # each key head serves r value heads
            q, k = q.repeat_interleave(r, dim=2), k.repeat_interleave(r, dim=2)
            v, z = v.reshape(b, l, hl * r, dv), z.reshape(l, b, hl * r, dv)
        beta = torch.sigmoid(beta_raw.reshape(l, b, hl * r).float()).permute(1, 0, 2)                                   # [b, l, h_v]
        g = (-torch.exp(self.A_log) * F.softplus(g_raw.reshape(l, b, hl * r).float() + self.dt_bias)).permute(1, 0, 2)  # log-decay ≤ 0
        if decode:
            o, S_new = gated_delta_rule_recurrent(q, k, v, g, beta, S)
            S.copy_(S_new)
        else:
            o, S_new = gated_delta_rule_chunked(q, k, v, g, beta, self.chunk_size)
            if inference_context is not None:
                inference_context.key_value_memory_dict[key] = (conv_state.clone(), S_new.clone())
        of = o.float()
        of = of * torch.rsqrt(of.pow(2).mean(-1, keepdim=True) + self.config.layernorm_epsilon) * self.norm_weight.float()
        of = of * F.silu(z.permute(1, 0, 2, 3).float())
        y = of.to(hidden_states.dtype).permute(1, 0, 2, 3).reshape(l, b, hl * r * dv)
        return self.out_proj(y)
