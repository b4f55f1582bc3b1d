"""Transformer helpers: sharded-state-dict builders, attention mask utils
(reference ``transformer/utils.py:96,252``)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import parallel_state as ps
from ..dist_checkpointing.mapping import ShardedObject, ShardedStateDict, StateDict
from ..utils import make_sharded_tensor_for_checkpoint, make_tp_sharded_tensor_for_checkpoint


def get_linear_layer(rows, columns, init_method, perform_initialization=True):
    layer = torch.nn.Linear(rows, columns)
    if perform_initialization:
        init_method(layer.weight)
    with torch.no_grad():
        layer.bias.zero_()
    return layer


def get_default_causal_mask(sq: int, device=None) -> torch.Tensor:
    return torch.triu(torch.ones(sq, sq, device=device), diagonal=1).bool()


def attention_mask_func(attention_scores, attention_mask):
    return attention_scores.masked_fill(attention_mask, -10000.0)


def make_sharded_tensors_for_checkpoint(
    state_dict: StateDict,
    prefix: str,
    tensor_parallel_layers_axis_map: Optional[Dict[str, int]] = None,
    sharded_offsets: Iterable[Tuple[int, int, int]] = (),
    extra_state_suffix: str = "_extra_state",
    tp_group=None,
    dp_cp_group=None,
) -> ShardedStateDict:
    """Wrap every tensor of ``state_dict``: keys in the axis map are TP-sharded along that
    axis, all others are TP-replicated; ``*_extra_state`` entries become ShardedObjects."""
    axis_map = tensor_parallel_layers_axis_map or {}
    out = {}
    for name, t in state_dict.items():
        key = f"{prefix}{name}"
        if name.endswith(extra_state_suffix):
            out[key] = make_sharded_object_for_checkpoint(t, key, sharded_offsets)
        elif name in axis_map:
            out[key] = make_tp_sharded_tensor_for_checkpoint(t, key, axis_map[name], prepend_offsets=sharded_offsets, tp_group=tp_group, dp_cp_group=dp_cp_group)
        else:
            out[key] = make_sharded_tensor_for_checkpoint(t, key, prepend_offsets=sharded_offsets, tp_group=tp_group, dp_cp_group=dp_cp_group)
    return out


def make_sharded_object_for_checkpoint(obj, key: str, sharded_offsets=(), replica_id=None, **kwargs):
    if replica_id is None:
        replica_id = (0, ps.get_tensor_model_parallel_rank(), ps.get_data_parallel_rank(with_context_parallel=True))
    return ShardedObject(key, obj, *_get_extra_state_offsets(sharded_offsets), replica_id, **kwargs)


def _get_extra_state_offsets(sharded_offsets):
    if sharded_offsets:
        so = sorted(sharded_offsets, key=lambda x: x[0])
        axis, off, shape = zip(*so)
        assert list(axis) == list(range(len(axis))), f"expected contiguous axis for offsets: {so}"
        return tuple(shape), tuple(off)
    return (1,), (0,)


def sharded_state_dict_default(module, prefix="", sharded_offsets=(), metadata=None, tp_group=None):
    """Use ``module.sharded_state_dict`` when it exists, else treat all tensors as replicated."""
    if hasattr(module, "sharded_state_dict"):
        return module.sharded_state_dict(prefix=prefix, sharded_offsets=sharded_offsets, metadata=metadata)
    sd = module.state_dict(prefix="", keep_vars=True)
    return make_sharded_tensors_for_checkpoint(sd, prefix, {}, sharded_offsets, tp_group=tp_group)


def ensure_metadata_has_dp_cp_group(metadata):
    metadata = dict(metadata or {})
    if "dp_cp_group" not in metadata:
        try:
            metadata["dp_cp_group"] = ps.get_data_parallel_group(with_context_parallel=True)
        except AssertionError:
            metadata["dp_cp_group"] = None
    return metadata


# ---- round-2 additions (reference ``transformer/utils.py``: masks, GELU variants, model-wide switches) ------------------
import math as _math


def get_sliding_window_causal_mask(sq: int, skv: int, window_size, device=None) -> torch.Tensor:
    """True = masked.  ``window_size = (left, right)``: query i (aligned to the END of the keys) sees keys in
    ``[i - left, i + right]``; ``-1`` means unbounded on that side.  The native attention kernels take the window itself
    (band mask, ``ops.attention_band``); this dense mask is for the unfused path and for tests."""
    left, right = window_size
    q = torch.arange(sq, device=device).unsqueeze(1) + (skv - sq)
    k = torch.arange(skv, device=device).unsqueeze(0)
    masked = torch.zeros(sq, skv, dtype=torch.bool, device=device)
    if left is not None and left >= 0:
        masked |= k < q - left
    if right is not None and right >= 0:
        masked |= k > q + right
    return masked


def is_layer_window_attention(window_size, window_attn_skip_freq, layer_number: int) -> bool:
    """Does (1-based) ``layer_number`` use sliding-window attention?  ``window_attn_skip_freq``: int N = every N-th layer is
    FULL attention; list = explicit 1/0 pattern per layer (1 = window)."""
    if window_size is None:
        return False
    if window_attn_skip_freq is None:
        return True
    if isinstance(window_attn_skip_freq, int):
        return layer_number % window_attn_skip_freq != 0
    return bool(window_attn_skip_freq[layer_number - 1])


def gelu_impl(x):
    """tanh-approximate GELU (the "OpenAI" GELU)."""
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def openai_gelu(x):
    return gelu_impl(x)


def erf_gelu(x):
    """Exact GELU written with erf (kept for checkpoints trained with it: identical to ``F.gelu`` up to rounding)."""
    return x * 0.5 * (torch.erf(x / _math.sqrt(2.0)).to(x.dtype) + torch.ones_like(x))


def cat_with_oom_fallback(tensors, dim: int = 0):
    """``torch.cat``; if the device allocator cannot serve the result, free the cache once, and as a last resort assemble
    the result on the host and move it back (used when gathering logits / KV for long sequences)."""
    try:
        return torch.cat(tensors, dim=dim)
    except torch.cuda.OutOfMemoryError:
        torch.cuda.empty_cache()
        try:
            return torch.cat(tensors, dim=dim)
        except torch.cuda.OutOfMemoryError:
            dev = tensors[0].device
            host = torch.cat([t.cpu() for t in tensors], dim=dim)
            del tensors
            torch.cuda.empty_cache()
            return host.to(dev)


def set_model_config_attribute(model, name: str, value) -> int:
    """Set ``config.<name>`` on every sub-module that has a config (configs may be shared or copied per layer)."""
    n, seen = 0, set()
    models = model if isinstance(model, (list, tuple)) else [model]
    for chunk in models:
        for m in chunk.modules():
            cfg = getattr(m, "config", None)
            if cfg is not None and id(cfg) not in seen and hasattr(cfg, name):
                setattr(cfg, name, value)
                seen.add(id(cfg))
                n += 1
    return n


def set_model_to_sequence_parallel(model, set_to: bool = False, exclude_modules=None) -> None:
    """Flip sequence parallelism on an already-built model (inference turns it off for single-token decode): the flag lives
    on the configs AND as an attribute on the TP linears / norms that captured it at construction."""
    exclude = set(exclude_modules or [])
    models = model if isinstance(model, (list, tuple)) else [model]
    for chunk in models:
        for name, m in chunk.named_modules():
            if any(name == e or name.startswith(e + ".") for e in exclude):
                continue
            if hasattr(m, "sequence_parallel"):
                m.sequence_parallel = set_to
            cfg = getattr(m, "config", None)
            if cfg is not None and hasattr(cfg, "sequence_parallel"):
                cfg.sequence_parallel = set_to
            for p in m.parameters(recurse=False):
                if hasattr(p, "sequence_parallel"):
                    p.sequence_parallel = set_to


def init_cuda_graph_cache(model) -> None:
    """Drop every captured graph of the model's graph managers (shapes / weights changed)."""
    models = model if isinstance(model, (list, tuple)) else [model]
    for chunk in models:
        for m in chunk.modules():
            mgr = getattr(m, "cudagraph_manager", None)
            if mgr is not None and hasattr(mgr, "captured"):
                mgr.captured.clear()


def toggle_cuda_graphs(model, set_to: str = "none", reset_cuda_graphs: bool = True) -> None:
    """Switch graph replay on/off model-wide between phases (RL: graphs for rollout decode, eager for training).
    ``set_to``: ``"none"`` | ``"local"`` (per-layer graphs) | ``"full"`` (whole-step capture handled by the trainer)."""
    assert set_to in ("none", "local", "full"), set_to
    models = model if isinstance(model, (list, tuple)) else [model]
    for chunk in models:
        for m in chunk.modules():
            mgr = getattr(m, "cudagraph_manager", None)
            if mgr is not None:
                mgr.enabled = set_to == "local"
            cfg = getattr(m, "config", None)
            if cfg is not None and hasattr(cfg, "cuda_graph_impl"):
                cfg.cuda_graph_impl = set_to
            if cfg is not None and hasattr(cfg, "enable_cuda_graph"):
                cfg.enable_cuda_graph = set_to == "local"
    if reset_cuda_graphs:
        init_cuda_graph_cache(model)


def transition_moe_cudagraphs(model, to_inference: bool) -> None:
    """MoE layers graph different regions in training (router + shared experts, dispatch eager) and in serving (the whole
    layer with static-shape dispatchers): drop the graphs of the other phase and record the phase on the layers."""
    models = model if isinstance(model, (list, tuple)) else [model]
    for chunk in models:
        for m in chunk.modules():
            if hasattr(m, "router") and hasattr(m, "experts"):
                m.cudagraph_phase = "inference" if to_inference else "training"
                mgr = getattr(m, "cudagraph_manager", None)
                if mgr is not None and hasattr(mgr, "captured"):
                    mgr.captured.clear()
