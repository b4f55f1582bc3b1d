"""MoE helpers: routing score functions, load-balancing losses, token permutation
(reference ``transformer/moe/moe_utils.py``: permute :344, unpermute :463, sort_chunks :672,
top-k routing :766-1017, losses, ``MoEAuxLossAutoScaler``)."""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.distributed as dist


# ---- losses -----------------------------------------------------------------------------------
def switch_load_balancing_loss_func(probs: torch.Tensor, tokens_per_expert: torch.Tensor, total_num_tokens: int, topk: int, num_experts: int,
                                    moe_aux_loss_coeff: float, fused: bool = False) -> torch.Tensor:
    """Switch-Transformer aux loss: ``E * coeff / (T^2 * k) * sum_e (sum_t p_te) * count_e``.  ``fused`` (reference: TE ``fused_moe_aux_loss``): one
    deterministic two-stage reduction kernel over the [T, E] probabilities and a broadcast backward (``ops/csrc/routing_kernels.cu``)."""
    if fused and probs.dim() == 2 and probs.is_cuda and probs.dtype == torch.float32:
        from .... import ops

        if ops.has_ext() and hasattr(ops.ext(), "moe_aux_loss_fwd"):
            return _FusedAuxLoss.apply(probs, tokens_per_expert, num_experts * moe_aux_loss_coeff / (topk * total_num_tokens * total_num_tokens))
    aggregated = probs.sum(dim=0) if probs.dim() == 2 else probs
    return torch.sum(aggregated * tokens_per_expert) * (num_experts * moe_aux_loss_coeff / (topk * total_num_tokens * total_num_tokens))


class _FusedAuxLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, tokens_per_expert, coeff):
        from .... import ops

        tpe = tokens_per_expert.float().contiguous()
        loss = ops.ext().moe_aux_loss_fwd(probs.contiguous(), tpe, float(coeff))
        ops._count(2)
        ctx.save_for_backward(tpe)
        ctx.coeff, ctx.T = float(coeff), probs.shape[0]
        return loss

    @staticmethod
    def backward(ctx, g):
        from .... import ops

        (tpe,) = ctx.saved_tensors
        gp = ops.ext().moe_aux_loss_bwd(tpe, g.float().contiguous(), ctx.coeff, ctx.T)
        ops._count()
        return gp, None, None


def z_loss_func(logits: torch.Tensor, z_loss_coeff: float) -> torch.Tensor:
    """ST-MoE router z-loss: mean(logsumexp(logits)^2) * coeff."""
    return torch.mean(torch.square(torch.logsumexp(logits.float(), dim=-1))) * z_loss_coeff


def sinkhorn(cost: torch.Tensor, tol: float = 1e-4) -> torch.Tensor:
    cost = torch.exp(cost)
    d0 = torch.ones(cost.size(0), device=cost.device, dtype=cost.dtype)
    d1 = torch.ones(cost.size(1), device=cost.device, dtype=cost.dtype)
    eps, error = 1e-8, 1e9
    d1_old = d1
    while error > tol:
        d0 = (1 / d0.size(0)) * 1 / (torch.sum(d1 * cost, 1) + eps)
        d1 = (1 / d1.size(0)) * 1 / (torch.sum(d0.unsqueeze(1) * cost, 0) + eps)
        error = torch.mean(torch.abs(d1_old - d1))
        d1_old = d1
    return d1 * cost * d0.unsqueeze(1)


def get_capacity(num_tokens: int, num_experts: int, capacity_factor: float, min_capacity: Optional[int] = None) -> int:
    cap = math.ceil((num_tokens / num_experts) * capacity_factor)
    if min_capacity is not None and cap < min_capacity:
        cap = min_capacity
    return cap


class MoEAuxLossAutoScaler(torch.autograd.Function):
    """Attach an auxiliary loss to an activation: forward is identity, backward injects
    ``d aux_loss = main_loss_backward_scale`` so the aux loss trains the router."""

    main_loss_backward_scale: Optional[torch.Tensor] = None

    @staticmethod
    def forward(ctx, output: torch.Tensor, aux_loss: torch.Tensor):
        ctx.save_for_backward(aux_loss)
        return output

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        (aux_loss,) = ctx.saved_tensors
        scale = MoEAuxLossAutoScaler.main_loss_backward_scale
        if scale is None:
            scale = torch.ones(1, device=aux_loss.device)
        return grad_output, torch.ones_like(aux_loss) * scale.to(aux_loss.device).reshape(()).to(aux_loss.dtype)

    @staticmethod
    def set_loss_scale(scale: torch.Tensor):
        MoEAuxLossAutoScaler.main_loss_backward_scale = scale.detach() if isinstance(scale, torch.Tensor) else torch.tensor(float(scale))


# ---- routing ------------------------------------------------------------------------------------
def group_limited_topk(scores: torch.Tensor, topk: int, num_tokens: int, num_experts: int, num_groups: int, group_topk: int):
    """DeepSeek device-limited routing: keep the best ``group_topk`` groups (by the sum of their
    top ``topk/group_topk`` scores), then take top-k inside them."""
    gs = scores.view(num_tokens, num_groups, -1).topk(max(topk // group_topk, 1), dim=-1)[0].sum(dim=-1)
    gidx = torch.topk(gs, k=group_topk, dim=-1, sorted=False)[1]
    gmask = torch.zeros_like(gs).scatter_(1, gidx, 1)
    smask = gmask.unsqueeze(-1).expand(num_tokens, num_groups, num_experts // num_groups).reshape(num_tokens, -1)
    masked = scores.masked_fill(~smask.bool(), float("-inf"))
    return torch.topk(masked, k=topk, dim=-1)


def topk_routing_with_score_function(logits: torch.Tensor, topk: int, use_pre_softmax: bool = False, num_groups: Optional[int] = None,
                                     group_topk: Optional[int] = None, scaling_factor: Optional[float] = None, score_function: str = "softmax",
                                     expert_bias: Optional[torch.Tensor] = None, fused: bool = False, router_replay=None):
    """→ (routing_probs [T, E] dense with zeros off the top-k, routing_map [T, E] bool).
    ``router_replay`` (``RouterReplay``) may record or override the selected indices."""
    assert logits.dim() == 2, f"expected 2D logits [num_tokens, num_experts], got {logits.dim()}"
    T, E = logits.shape

    def _pick(scores, k):
        if num_groups:
            return group_limited_topk(scores, k, T, E, num_groups, group_topk)
        return torch.topk(scores, k=k, dim=1)

    def pick(scores, k):
        if router_replay is not None:
            return router_replay.get_replay_topk(scores, k, _pick)
        return _pick(scores, k)

    if score_function == "softmax":
        if use_pre_softmax:
            scores = torch.softmax(logits, dim=-1, dtype=torch.float32).type_as(logits)
            probs, idx = pick(scores, topk)
        else:
            scores, idx = pick(logits, topk)
            probs = torch.softmax(scores, dim=-1, dtype=torch.float32).type_as(logits)
    elif score_function in ("sigmoid", "sqrtsoftplus"):
        scores = torch.sigmoid(logits.float()).type_as(logits) if score_function == "sigmoid" else torch.sqrt(torch.nn.functional.softplus(logits.float())).type_as(logits)
        if expert_bias is not None:
            _, idx = pick(scores + expert_bias, topk)
            scores = torch.gather(scores, dim=1, index=idx).type_as(logits)
        else:
            scores, idx = pick(scores, topk)
        probs = scores / (scores.sum(dim=-1, keepdim=True) + 1e-20) if topk > 1 else scores
    else:
        raise ValueError(f"invalid score_function: {score_function}")
    if scaling_factor:
        probs = probs * scaling_factor
    routing_probs = torch.zeros_like(logits).scatter(1, idx, probs)
    routing_map = torch.zeros_like(logits).int().scatter(1, idx, 1).bool()
    return routing_probs, routing_map


def apply_random_logits(logits):
    """Force-balanced routing for benchmarking (``moe_router_force_load_balancing``)."""
    return _RandomSTE.apply(logits)


class _RandomSTE(torch.autograd.Function):
    generator = None

    @staticmethod
    def forward(ctx, logits):
        if _RandomSTE.generator is None or _RandomSTE.generator.device != logits.device:
            _RandomSTE.generator = torch.Generator(device=logits.device)
            _RandomSTE.generator.manual_seed(1234 + (dist.get_rank() if dist.is_initialized() else 0))
        return logits.clone().normal_(generator=_RandomSTE.generator)

    @staticmethod
    def backward(ctx, g):
        return g


def apply_router_token_dropping(routing_probs, routing_map, router_topk, capacity_factor, drop_policy="probs", pad_to_capacity=False):
    """Cap every expert at ``capacity`` tokens, dropping by lowest prob or by position."""
    T, E = routing_probs.shape
    cap = get_capacity(T * router_topk, E, capacity_factor)
    if drop_policy == "probs":
        _, cidx = torch.topk(routing_probs, k=min(cap, T), dim=0, sorted=False)
        cmask = torch.zeros_like(routing_probs).scatter(0, cidx, 1).bool()
    elif drop_policy == "position":
        _, cidx = torch.topk(routing_map.int(), k=min(cap, T), dim=0, sorted=False)
        cmask = torch.zeros_like(routing_probs).scatter(0, cidx, 1).bool()
    else:
        raise ValueError(f"invalid drop_policy: {drop_policy}")
    if pad_to_capacity:
        final_map = cmask
        final_probs = routing_probs * final_map
    else:
        final_map = torch.logical_and(routing_map, cmask)
        final_probs = routing_probs * final_map
    return final_probs, final_map


# ---- permutation -----------------------------------------------------------------------------------
def permute(tokens: torch.Tensor, routing_map: torch.Tensor, probs: Optional[torch.Tensor] = None, num_out_tokens: Optional[int] = None,
            fused: bool = False, drop_and_pad: bool = False):
    """Gather token copies grouped by expert.

    ``routing_map`` [T, E] bool.  Output row order: all tokens of expert 0 (in token order),
    then expert 1, …  Returns ``(permuted_tokens, permuted_probs, sorted_indices)`` where
    ``sorted_indices[i]`` is the source token of output row i."""
    T, E = routing_map.shape
    rm = routing_map.bool().T.contiguous()  # [E, T]
    if drop_and_pad and num_out_tokens is not None:
        cap = num_out_tokens // E
        sorted_indices = rm.to(torch.int8).argsort(dim=-1, descending=True, stable=True)[:, :cap].contiguous().view(-1)
        permuted_probs = None
        if probs is not None:
            pT = probs.T.contiguous()
            permuted_probs = pT.gather(1, sorted_indices.view(E, cap)).reshape(-1)
    else:
        token_idx = torch.arange(T, device=routing_map.device).unsqueeze(0).expand(E, -1)
        sorted_indices = token_idx.masked_select(rm)
        permuted_probs = probs.T.contiguous().masked_select(rm) if probs is not None else None
    return tokens.index_select(0, sorted_indices), permuted_probs, sorted_indices


def unpermute(permuted_tokens: torch.Tensor, sorted_indices: torch.Tensor, restore_shape: torch.Size, probs: Optional[torch.Tensor] = None,
              routing_map: Optional[torch.Tensor] = None, fused: bool = False, drop_and_pad: bool = False):
    """Scatter-add expert outputs back to token order (optionally weighting by ``probs``)."""
    T, hidden = restore_shape
    if probs is not None:
        assert routing_map is not None, "unpermute with probs needs the routing map"
        if drop_and_pad:
            E = routing_map.size(1)
            cap = sorted_indices.numel() // E
            pp = probs.T.contiguous().gather(1, sorted_indices.view(E, cap)).reshape(-1)
        else:
            pp = probs.T.contiguous().masked_select(routing_map.bool().T.contiguous())
        permuted_tokens = permuted_tokens * pp.unsqueeze(-1).to(permuted_tokens.dtype)
    out = torch.zeros(restore_shape, dtype=permuted_tokens.dtype, device=permuted_tokens.device)
    out.index_add_(0, sorted_indices, permuted_tokens)
    return out


def sort_chunks_by_idxs(inp: torch.Tensor, split_sizes, sorted_idxs, probs: Optional[torch.Tensor] = None, fused: bool = False):
    """Reorder contiguous chunks of rows: chunk i has ``split_sizes[i]`` rows; output = chunks in ``sorted_idxs`` order."""
    ss = split_sizes.tolist() if isinstance(split_sizes, torch.Tensor) else list(split_sizes)
    idx = sorted_idxs.tolist() if isinstance(sorted_idxs, torch.Tensor) else list(sorted_idxs)
    chunks = torch.split(inp, ss, dim=0)
    out = torch.cat([chunks[i] for i in idx], dim=0)
    if probs is not None:
        pc = torch.split(probs, ss, dim=0)
        return out, torch.cat([pc[i] for i in idx], dim=0)
    return out, None


# ---- logging tracker ----------------------------------------------------------------------------------
_MOE_LAYER_WISE_LOGGING_TRACKER = {}


def save_to_aux_losses_tracker(name: str, loss: torch.Tensor, layer_number: int, num_layers: int, reduce_group=None, avg_group=None):
    if layer_number is None:
        return
    t = _MOE_LAYER_WISE_LOGGING_TRACKER
    if name not in t:
        t[name] = {"values": torch.zeros(num_layers, device=loss.device), "reduce_group": None, "avg_group": None}
    t[name]["values"][layer_number - 1] += loss.detach()
    t[name]["reduce_group"], t[name]["avg_group"] = reduce_group, avg_group


def get_moe_layer_wise_logging_tracker():
    return _MOE_LAYER_WISE_LOGGING_TRACKER


def clear_aux_losses_tracker():
    for v in _MOE_LAYER_WISE_LOGGING_TRACKER.values():
        v["values"].zero_()


def reduce_aux_losses_tracker_across_ranks(track_names: Optional[List[str]] = None):
    from ... import parallel_state as ps

    for name, rec in _MOE_LAYER_WISE_LOGGING_TRACKER.items():
        if track_names is not None and name not in track_names:
            continue
        v = rec["values"]
        if ps.is_initialized() and ps.get_pipeline_model_parallel_world_size() > 1:
            dist.all_reduce(v, group=ps.get_pipeline_model_parallel_group())
        if rec.get("reduce_group") is not None:
            dist.all_reduce(v, group=rec["reduce_group"])
        if rec.get("avg_group") is not None:
            dist.all_reduce(v, group=rec["avg_group"])
            v.div_(dist.get_world_size(rec["avg_group"]))


def track_moe_metrics(loss_scale, iteration, writer=None, wandb_writer=None, total_loss_dict=None, per_layer_logging=False, force_initialize=False,
                      track_names=None, num_layers=None, moe_layer_freq=None, mtp_num_layers=None):
    reduce_aux_losses_tracker_across_ranks(track_names)
    out = {}
    for name, rec in _MOE_LAYER_WISE_LOGGING_TRACKER.items():
        vals = rec["values"].float() * loss_scale
        n = max(int((vals != 0).sum()), 1)
        out[name] = float(vals.sum() / n)
        if total_loss_dict is not None:
            total_loss_dict[name] = total_loss_dict.get(name, 0.0) + out[name]
        if writer is not None:
            writer.add_scalar(name, out[name], iteration)
            if per_layer_logging:
                for i, lv in enumerate(vals.tolist()):
                    writer.add_scalar(f"moe/{name}_layer_{i}", lv, iteration)
    clear_aux_losses_tracker()
    return out


# ---- round-2 additions: helpers the router / dispatchers expose by name (reference ``moe_utils.py``) ---------------------
RandomSTE = _RandomSTE


def get_tokens_per_expert_and_token_count(routing_map: torch.Tensor, reduce_group=None, topk: Optional[int] = None, with_padding_mask: bool = False):
    """(global tokens per expert, local token count, global token count).  With a padding mask the routing map has all-zero rows
    for padding tokens, so the counts come from the map itself (``sum / topk``) instead of its height."""
    local = routing_map.sum(dim=0)
    world = torch.distributed.get_world_size(reduce_group) if (reduce_group is not None and torch.distributed.is_initialized()) else 1
    glob = local.clone()
    if world > 1:
        torch.distributed.all_reduce(glob, group=reduce_group)
    if with_padding_mask:
        assert topk, "topk is needed to count tokens under a padding mask"
        return glob, local.sum() / topk, glob.sum() / topk
    return glob, routing_map.shape[0], routing_map.shape[0] * world


def compute_routing_scores_for_aux_loss(logits: torch.Tensor, topk: int, score_function: str, fused: bool = False, padding_mask: Optional[torch.Tensor] = None):
    """Scores used by the balancing losses: NORMALISED over all experts (softmax, or sigmoid / sqrt-softplus divided by their
    row sum) with the plain top-k map of those scores — independent of the bias / group limits the real routing applies.
    ``padding_mask``: True = padding token (contributes nothing).  ``fused`` selects the CUDA top-k kernel on GPU."""
    x = logits.float()
    if score_function == "softmax":
        scores = torch.softmax(x, dim=-1)
    elif score_function in ("sigmoid", "sqrtsoftplus"):
        scores = torch.sigmoid(x) if score_function == "sigmoid" else torch.nn.functional.softplus(x).sqrt()
        scores = scores / (scores.sum(dim=-1, keepdim=True) + 1e-20)
    else:
        raise ValueError(f"Invalid score_function: {score_function}")
    top = torch.topk(scores, k=topk, dim=1).indices
    routing_map = torch.zeros_like(logits, dtype=torch.bool).scatter_(1, top, True)
    if padding_mask is not None:
        keep = (~padding_mask).unsqueeze(-1)
        routing_map, scores = routing_map & keep, scores * keep
    return routing_map, scores


def get_updated_expert_bias(tokens_per_expert: torch.Tensor, expert_bias: torch.Tensor, expert_bias_update_rate: float, tp_dp_cp_group=None) -> torch.Tensor:
    """Aux-loss-free balancing (arXiv 2408.15664): experts that received fewer tokens than the average get their routing bias
    raised by ``rate``, the others lowered.  Counts are summed over every rank that sees different tokens of the step.
    Accepts stacked [num_layers, num_experts] inputs so that ONE all-reduce serves all layers."""
    with torch.no_grad():
        if tp_dp_cp_group is None and torch.distributed.is_initialized():
            from ... import parallel_state as ps
            tp_dp_cp_group = ps.get_tensor_and_data_parallel_group(with_context_parallel=True) if ps.model_parallel_is_initialized() else None
        if tp_dp_cp_group is not None and torch.distributed.get_world_size(tp_dp_cp_group) > 1:
            torch.distributed.all_reduce(tokens_per_expert, group=tp_dp_cp_group)
        mean = tokens_per_expert.sum(dim=-1, keepdim=True) / tokens_per_expert.shape[-1]
        return expert_bias + torch.sign(mean - tokens_per_expert) * expert_bias_update_rate


def maybe_move_tensor_to_cpu(tensor, as_numpy: bool = False, record_stream: bool = False):
    """Non-blocking device->pinned-host copy of small routing statistics (tokens per expert); ``record_stream`` keeps the
    source alive until the copy has run when the caller drops it immediately."""
    if torch.is_tensor(tensor) and tensor.is_cuda:
        host = torch.empty(tensor.shape, dtype=tensor.dtype, device="cpu", pin_memory=True)
        host.copy_(tensor, non_blocking=True)
        if record_stream:
            tensor.record_stream(torch.cuda.current_stream())
        tensor = host
    return tensor.numpy() if (as_numpy and torch.is_tensor(tensor)) else tensor


class RouterGatingLinearFunction(torch.autograd.Function):
    """Gating GEMM in ``router_dtype`` (fp32 / fp64) without keeping up-cast copies of the activations: inputs are saved in
    their own dtype and re-cast in backward; gradients come back in the inputs' dtypes."""

    @staticmethod
    def forward(ctx, inp, weight, bias, router_dtype):
        ctx.save_for_backward(inp, weight, bias)
        ctx.router_dtype = router_dtype
        x = inp.reshape(-1, inp.shape[-1]).to(router_dtype)
        out = x @ weight.to(router_dtype).t()
        if bias is not None:
            out = out + bias.to(router_dtype)
        return out.view(*inp.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        inp, weight, bias = ctx.saved_tensors
        dt = ctx.router_dtype
        g2 = g.reshape(-1, g.shape[-1]).to(dt)
        gi = (g2 @ weight.to(dt)).to(inp.dtype).view(inp.shape) if ctx.needs_input_grad[0] else None
        gw = (g2.t() @ inp.reshape(-1, inp.shape[-1]).to(dt)).to(weight.dtype) if ctx.needs_input_grad[1] else None
        gb = g2.sum(0).to(bias.dtype) if (bias is not None and ctx.needs_input_grad[2]) else None
        return gi, gw, gb, None


def router_gating_linear(inp: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], router_dtype: torch.dtype):
    return RouterGatingLinearFunction.apply(inp, weight, bias, router_dtype)


def get_align_size_for_quantization(config) -> int:
    """Multiple every expert's token count is padded to so that the block-scaled grouped GEMM sees whole scale blocks:
    128 rows (one scale atom of the tcgen05 block-scaled MMA) for MXFP8 / NVFP4, 16 for per-tensor FP8, none for bf16."""
    if getattr(config, "fp4", None):
        return 128
    if getattr(config, "fp8", None):
        return 128 if getattr(config, "fp8_recipe", None) in ("mxfp8", "blockwise") else 16
    return 0


def get_default_pg_collection():
    """The process groups MoE layers use when none are passed: read from the global parallel state."""
    from ...process_groups_config import ProcessGroupCollection
    return ProcessGroupCollection.use_mpu_process_groups()
