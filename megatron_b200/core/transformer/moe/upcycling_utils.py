"""Dense → MoE "upcycling" of a trained checkpoint (reference ``transformer/moe/upcycling_utils.py:16-358``;
granular variant after arXiv:2410.07524).

Every dense FFN becomes ``E = granularity × expansion`` experts.  The FFN dimension is cut into
``granularity`` shards; expert ``e`` receives shard ``e % granularity`` and experts
``[c·G, (c+1)·G)`` (one complete copy ``c`` of the dense FFN) share one router row, so a token
that picks a copy picks all of its shards.  Because top-k probabilities sum to one, a copy's
shards arrive weighted ``1/G`` of what the dense layer computed; the weights are therefore
scaled so that the product through the (roughly degree-``d`` homogeneous) FFN restores it:
``s = scale^(1/d)``, ``d = 3`` for gated or squared-ReLU FFNs and ``2`` otherwise.

With ``expansion = 1`` and ``topk = granularity`` (and a degree-1 activation) the upcycled layer
reproduces the dense layer exactly, which is what the unit test checks.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Tuple

import torch

from ...utils import get_pg_rank


def _get_keys_endswith(state_dict, suffix: str) -> List[str]:
    return [k for k in state_dict if k.endswith(suffix)]


def _find_submodule(model, name: str):
    for n, m in model.named_modules():
        if n.endswith(name):
            return m
    return None


def _get_config(moe_model, dense_model) -> Tuple[int, int, int, int, str, bool, str, bool, int]:
    mc, dc = moe_model.config, dense_model.config
    dense_ffn, moe_ffn = dc.ffn_hidden_size, mc.moe_ffn_hidden_size or mc.ffn_hidden_size
    if dense_ffn % moe_ffn:
        raise ValueError(f"dense ffn {dense_ffn} is not a multiple of the expert ffn {moe_ffn}")
    granularity = dense_ffn // moe_ffn
    if mc.num_moe_experts % granularity:
        raise ValueError(f"{mc.num_moe_experts} experts cannot hold whole copies of {granularity} shards")
    expansion = mc.num_moe_experts // granularity
    experts = _find_submodule(moe_model, "mlp.experts")
    experts_type = "grouped" if hasattr(experts, "weight1") else "sequential"
    ep_rank = get_pg_rank(getattr(experts, "ep_group", None))
    act = getattr(mc.activation_func, "__name__", str(mc.activation_func))
    return experts.num_local_experts, mc.moe_router_topk, granularity, expansion, experts_type, bool(mc.gated_linear_unit), act, bool(mc.moe_router_pre_softmax), ep_rank


def _weight_scale(granularity, expansion, topk, gated, act, pre_softmax) -> float:
    out_scale = (expansion * granularity * granularity) / topk if pre_softmax else float(granularity)
    degree = 3 if (gated or act in ("squared_relu", "relu2")) else 2
    return out_scale ** (1.0 / degree)


def _convert_to_moe_state_dict(moe_model, dense_model) -> Dict[str, torch.Tensor]:
    n_local, topk, G, X, experts_type, gated, act, pre_softmax, ep_rank = _get_config(moe_model, dense_model)
    s = _weight_scale(G, X, topk, gated, act, pre_softmax)
    dense_sd = dense_model.state_dict()
    moe_sd = copy.deepcopy(moe_model.state_dict())
    # everything that is not the FFN (attention, norms, embeddings) carries over unchanged
    for k in dense_sd.keys() & moe_sd.keys():
        if torch.is_tensor(dense_sd[k]):
            moe_sd[k] = dense_sd[k].clone()

    def fc1_shards(w):
        """``[gate‖up, H]`` or ``[F, H]`` → G tensors holding the matching rows of each half."""
        w = w * s
        if gated:
            a, b = torch.chunk(w, 2, dim=0)
            return [torch.cat([x, y], 0) for x, y in zip(torch.tensor_split(a, G, 0), torch.tensor_split(b, G, 0))]
        return list(torch.tensor_split(w, G, 0))

    def fc2_shards(w):
        return list(torch.tensor_split(w * s, G, dim=1))

    first = ep_rank * n_local
    for k1 in _get_keys_endswith(dense_sd, "mlp.linear_fc1.weight"):
        pre = k1[: -len("linear_fc1.weight")]                        # "...layers.N.mlp."
        if not any(k.startswith(pre + "experts.") for k in moe_sd):
            continue                                                   # this layer stays dense (moe_layer_freq)
        w1, w2 = fc1_shards(dense_sd[k1]), fc2_shards(dense_sd[pre + "linear_fc2.weight"])
        b1 = dense_sd.get(pre + "linear_fc1.bias")
        b2 = dense_sd.get(pre + "linear_fc2.bias")
        b1s = fc1_shards(b1.unsqueeze(-1)) if b1 is not None else None
        for j in range(n_local):
            shard = (first + j) % G
            if experts_type == "grouped":
                moe_sd[pre + "experts.weight1"][j].copy_(w1[shard])
                moe_sd[pre + "experts.weight2"][j].copy_(w2[shard])
            else:
                e = f"{pre}experts.local_experts.{j}."
                moe_sd[e + "linear_fc1.weight"] = w1[shard].clone()
                moe_sd[e + "linear_fc2.weight"] = w2[shard].clone()
                if b1s is not None and e + "linear_fc1.bias" in moe_sd:
                    moe_sd[e + "linear_fc1.bias"] = b1s[shard].squeeze(-1).clone()
                if b2 is not None and e + "linear_fc2.bias" in moe_sd:
                    # the G shards of one copy each add their bias with weight 1/G·(out scale) → keep the dense bias
                    moe_sd[e + "linear_fc2.bias"] = b2.clone()
        rk = pre + "router.weight"
        r = moe_sd[rk]
        rows = r[: r.shape[0] // G]                                    # one row per copy
        moe_sd[rk] = rows.repeat_interleave(G, dim=0).contiguous()
    return moe_sd


def upcycle_state_dict(moe_model: List[torch.nn.Module], dense_model: List[torch.nn.Module]) -> Dict[str, Dict[str, torch.Tensor]]:
    """Model chunks in, ``{"model": sd}`` (or ``model0``, ``model1``, … for virtual PP) out."""
    moe_model = moe_model if isinstance(moe_model, (list, tuple)) else [moe_model]
    dense_model = dense_model if isinstance(dense_model, (list, tuple)) else [dense_model]
    if len(moe_model) != len(dense_model):
        raise ValueError("dense and MoE models must have the same number of chunks")
    if len(moe_model) == 1:
        return {"model": _convert_to_moe_state_dict(moe_model[0], dense_model[0])}
    return {f"model{i}": _convert_to_moe_state_dict(m, d) for i, (m, d) in enumerate(zip(moe_model, dense_model))}


def load_and_upcycle_model(load_dense_ckpt_func, moe_model, dense_model, strict: bool = True, load_args=(), load_kwargs=None):
    """Load the dense checkpoint with the caller's loader, then fill the MoE chunks from it."""
    iteration, flops = load_dense_ckpt_func(*load_args, **(load_kwargs or {}))
    sd = upcycle_state_dict(moe_model, dense_model)
    chunks = moe_model if isinstance(moe_model, (list, tuple)) else [moe_model]
    for i, m in enumerate(chunks):
        m.load_state_dict(sd["model" if len(chunks) == 1 else f"model{i}"], strict=strict)
    return iteration, flops
