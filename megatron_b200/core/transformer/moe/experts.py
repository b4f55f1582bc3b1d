"""Expert networks (reference ``transformer/moe/experts.py``: ``TEGroupedMLP`` :183, ``SequentialMLP`` :1409).

``GroupedMLP`` stores all local experts' weights in ONE tensor per projection
(``weight1 [L, 2F, H]``, ``weight2 [L, H, F]``) and runs the grouped GEMM
(``megatron_b200.ops.grouped``: one launch over all experts' token groups); ``SequentialMLP``
is the per-expert loop kept for checkpoint/layout parity.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .... import ops
from ... import parallel_state as ps
from ...dist_checkpointing.mapping import ShardedTensor
from ...tensor_parallel.layers import set_tensor_model_parallel_attributes
from ...tensor_parallel.random import get_cuda_rng_tracker, get_expert_parallel_rng_tracker_name
from ...utils import divide, get_pg_rank, get_pg_size
from ..mlp import MLP, MLPSubmodules
from ..module import MegatronModule
from ..transformer_config import TransformerConfig


def _expert_groups(pg_collection):
    if pg_collection is not None:
        return pg_collection.ep, pg_collection.expt_tp, getattr(pg_collection, "expt_dp", None)
    return (ps.get_expert_model_parallel_group(check_initialized=False), ps.get_expert_tensor_parallel_group(check_initialized=False),
            ps.get_group("expt_dp", check_initialized=False))


class _GroupedLinearFn(torch.autograd.Function):
    """``out[rows of expert e] = x[rows of expert e] @ W[e]ᵀ`` for all local experts."""

    @staticmethod
    def forward(ctx, x, w, tokens_per_expert, accumulate_main_grad):
        from ....ops import grouped

        ctx.save_for_backward(x, w)
        ctx.tpe, ctx.acc = tokens_per_expert, accumulate_main_grad
        return grouped.grouped_gemm_nt(x, w, tokens_per_expert)

    @staticmethod
    def backward(ctx, gy):
        from ....ops import grouped

        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = grouped.grouped_gemm_nn(gy, w, ctx.tpe)
        gw = grouped.grouped_gemm_tn(gy, x, ctx.tpe, w)
        return gx, gw, None, None


class GroupedMLP(MegatronModule):
    def __init__(self, num_local_experts: int, config: TransformerConfig, submodules: Optional[MLPSubmodules] = None, pg_collection=None):
        super().__init__(config)
        if config.add_bias_linear:
            raise ValueError("the grouped-GEMM experts have no bias terms: use add_bias_linear=False (--disable-bias-linear) or moe_grouped_gemm=False")
        self.num_local_experts = num_local_experts
        self.ep_group, self.tp_group, self.expt_dp_group = _expert_groups(pg_collection)
        tp = get_pg_size(self.tp_group)
        F_ = divide(config.moe_ffn_hidden_size, tp)
        self.ffn_per_partition = F_
        H = config.hidden_size
        gate = 2 if config.gated_linear_unit else 1
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight1 = torch.nn.Parameter(torch.empty(num_local_experts, gate * F_, H, device=dev, dtype=config.params_dtype))
        self.weight2 = torch.nn.Parameter(torch.empty(num_local_experts, H, F_, device=dev, dtype=config.params_dtype))
        if config.perform_initialization:
            tracker = get_cuda_rng_tracker()
            ctx = tracker.fork(get_expert_parallel_rng_tracker_name()) if tracker.is_initialized() else torch.no_grad()
            with ctx:
                for e in range(num_local_experts):
                    config.init_method(self.weight1.data[e])
                    config.output_layer_init_method(self.weight2.data[e])
        set_tensor_model_parallel_attributes(self.weight1, True, 1, 1)
        set_tensor_model_parallel_attributes(self.weight2, True, 2, 1)
        ep_on = get_pg_size(self.ep_group) > 1
        setattr(self.weight1, "allreduce", not ep_on)
        setattr(self.weight2, "allreduce", not ep_on)
        self.activation_func = config.activation_func

    def forward(self, permuted_tokens: torch.Tensor, tokens_per_expert, permuted_probs: Optional[torch.Tensor] = None):
        """tokens are grouped by local expert; ``tokens_per_expert`` [L] (host tensor/list)."""
        tpe = tokens_per_expert.tolist() if isinstance(tokens_per_expert, torch.Tensor) else list(tokens_per_expert)
        probs = permuted_probs.unsqueeze(-1) if permuted_probs is not None else None
        if self.config.moe_apply_probs_on_input and probs is not None:
            permuted_tokens = (permuted_tokens * probs.to(permuted_tokens.dtype))
            probs = None
        if permuted_tokens.shape[0] == 0:
            # keep the graph connected so every rank produces expert grads
            z = (self.weight1.sum() + self.weight2.sum()) * 0.0
            return permuted_tokens + z.to(permuted_tokens.dtype), None
        acc = self.config.gradient_accumulation_fusion
        h1 = _GroupedLinearFn.apply(permuted_tokens, self.weight1, tpe, acc)
        if self.config.gated_linear_unit and self.activation_func is F.silu:
            act = ops.swiglu(h1, None, probs.float() if probs is not None else None)
        else:
            if self.config.gated_linear_unit:
                a, b = h1.chunk(2, dim=-1)
                act = self.activation_func(a) * b
            else:
                act = self.activation_func(h1)
            if probs is not None:
                act = act * probs.to(act.dtype)
        out = _GroupedLinearFn.apply(act, self.weight2, tpe, acc)
        return out, None

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: Optional[dict] = None):
        """On disk: ``…experts.experts.linear_fc{1,2}.weight`` with a leading global-expert axis, so EP
        re-partitioning is free (reference :1049-1091)."""
        ep, epr = get_pg_size(self.ep_group), get_pg_rank(self.ep_group)
        tp, tpr = get_pg_size(self.tp_group), get_pg_rank(self.tp_group)
        L = self.num_local_experts
        E = L * ep
        dp_rank = get_pg_rank(self.expt_dp_group) if self.expt_dp_group is not None else 0
        rid = (0, 0, dp_rank)
        out = {}
        gate = 2 if self.config.gated_linear_unit else 1
        n = len(sharded_offsets)
        w1 = self.weight1
        if gate == 2:
            # split [gate; up] so each half is TP-sharded along its own axis
            g, u = w1.chunk(2, dim=1)
            for nm, t, off in (("gate", g, tpr), ("up", u, tp + tpr)):
                out[f"{prefix}weight1.{nm}"] = ShardedTensor.from_rank_offsets(
                    f"{prefix}experts.linear_fc1.weight", t.contiguous(), *sharded_offsets, (n + 0, epr, ep), (n + 1, off, 2 * tp), replica_id=rid, prepend_axis_num=n)
        else:
            out[f"{prefix}weight1"] = ShardedTensor.from_rank_offsets(
                f"{prefix}experts.linear_fc1.weight", w1, *sharded_offsets, (n + 0, epr, ep), (n + 1, tpr, tp), replica_id=rid, prepend_axis_num=n)
        out[f"{prefix}weight2"] = ShardedTensor.from_rank_offsets(
            f"{prefix}experts.linear_fc2.weight", self.weight2, *sharded_offsets, (n + 0, epr, ep), (n + 2, tpr, tp), replica_id=rid, prepend_axis_num=n)
        return out


class SequentialMLP(MegatronModule):
    """One ``MLP`` per local expert, executed in a loop over the token groups."""

    def __init__(self, num_local_experts: int, config: TransformerConfig, submodules: MLPSubmodules, pg_collection=None):
        super().__init__(config)
        self.add_bias = config.add_bias_linear
        self.num_local_experts = num_local_experts
        self.ep_group, self.tp_group, self.expt_dp_group = _expert_groups(pg_collection)
        self.local_experts = torch.nn.ModuleList(
            [MLP(config, submodules, ffn_hidden_size=config.moe_ffn_hidden_size, is_expert=True, tp_group=self.tp_group) for _ in range(num_local_experts)]
        )

    def forward(self, permuted_tokens: torch.Tensor, tokens_per_expert, permuted_probs: Optional[torch.Tensor] = None):
        tpe = tokens_per_expert.tolist() if isinstance(tokens_per_expert, torch.Tensor) else list(tokens_per_expert)
        probs = permuted_probs
        if self.config.moe_apply_probs_on_input and probs is not None:
            permuted_tokens = permuted_tokens * probs.unsqueeze(-1).to(permuted_tokens.dtype)
            probs = None
        chunks = torch.split(permuted_tokens, tpe)
        pchunks = torch.split(probs, tpe) if probs is not None else [None] * len(tpe)
        outs = []
        for expert, x, p in zip(self.local_experts, chunks, pchunks):
            o, b = expert(x, per_token_scale=p.unsqueeze(-1) if p is not None else None)
            if self.add_bias and b is not None:
                # the output bias of an expert belongs to the tokens routed to it, weighted like the rest of its output: fold it in here — the
                # combine step only knows about one tensor (the reference asserts ``mlp_bias is None`` after the experts, moe_layer.py:568)
                o = o + (b * p.unsqueeze(-1).to(b.dtype) if p is not None else b)
            outs.append(o)
        return torch.cat(outs, dim=0), None

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: tuple = (), metadata: Optional[dict] = None):
        ep, epr = get_pg_size(self.ep_group), get_pg_rank(self.ep_group)
        E = self.num_local_experts * ep
        out = {}
        dp_rank = get_pg_rank(self.expt_dp_group) if self.expt_dp_group is not None else 0
        for i, expert in enumerate(self.local_experts):
            g = epr * self.num_local_experts + i
            sd = expert.sharded_state_dict(f"{prefix}local_experts.{i}.", sharded_offsets, metadata)
            for k, v in sd.items():
                def fix(t):
                    if hasattr(t, "key"):
                        t.key = t.key.replace(f"{prefix}local_experts.{i}.", f"{prefix}experts.{g}.")
                        if hasattr(t, "replica_id") and isinstance(t.replica_id, tuple):
                            t.replica_id = (*t.replica_id[:2], dp_rank)
                    return t

                if hasattr(v, "build_fn"):
                    inner = v.build_fn

                    def wrapped(key, t, rid, fr, _inner=inner, _old=f"{prefix}local_experts.{i}.", _new=f"{prefix}experts.{g}."):
                        return [fix(x) for x in _inner(key.replace(_old, _new), t, rid, fr)]

                    v.build_fn = wrapped
                    v.key = v.key.replace(f"{prefix}local_experts.{i}.", f"{prefix}experts.{g}.")
                else:
                    fix(v)
                out[k] = v
        return out
