"""DeepSeek sparse attention (DSA, DeepSeek-V3.2) — reference ``experimental_attention_variant/dsa.py`` (+ ``dsa_indexer_loss.py``, ``dsa_masking.py``).

A light *indexer* scores every (query, key) pair with a few low-dimensional heads,

    I[t, s] = Σ_h  w[t, h] · relu( q_idx[t, h] · k_idx[s] ),

keeps the ``index_topk`` best causal keys per query, and the main attention (absorbed MLA: one latent key/value row per token) runs on those keys
only: O(s·k) instead of O(s²).  The indexer is trained by a KL term that pulls softmax(I) towards the head-averaged attention distribution of the
main branch (target detached); ``DSAIndexerLossAutoScaler`` attaches that loss to the activations like the MoE aux loss.

Kernels: the selection is a ``topk`` over the score matrix and the sparse attention a gather of ``k`` latent rows per query followed by a batched
[1 x k] · [k x (r + d_rope)] product — bandwidth-bound on the gathered rows, which is why it is formulated on the absorbed latent (576 elements per
key instead of 40 960).  This module implements both steps with PyTorch ops on the gathered tensors; the layers to skip re-selection
(``topk_freq`` / ``skip_topk_offset``) reuse the indices of the layer that computed them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch

from ..module import MegatronModule
from ..spec_utils import ModuleSpec, build_module


def is_dsa_skip_topk_layer(layer_number: int, skip_topk_offset: int, topk_freq: int) -> bool:
    """Layers (1-based) that REUSE the selection of an earlier layer: with ``topk_freq = f`` only every f-th layer after the offset runs the indexer."""
    if topk_freq <= 1 or layer_number <= skip_topk_offset:
        return False
    return (layer_number - skip_topk_offset - 1) % topk_freq != 0


def source_dsa_compute_layer(layer_number: int, skip_topk_offset: int, topk_freq: int) -> int:
    """The layer whose indices ``layer_number`` uses (itself when it computes them)."""
    if not is_dsa_skip_topk_layer(layer_number, skip_topk_offset, topk_freq):
        return layer_number
    return layer_number - (layer_number - skip_topk_offset - 1) % topk_freq


def rotate_activation(x: torch.Tensor) -> torch.Tensor:
    """Hadamard rotation of the last dim (power of two), scaled to be orthonormal — spreads outliers before the low-precision index dot products."""
    d = x.shape[-1]
    assert d & (d - 1) == 0, "index head dim must be a power of two"
    y = x.float().reshape(-1, d)
    h = 1
    while h < d:
        y = y.view(-1, d // (2 * h), 2, h)
        y = torch.stack([y[:, :, 0] + y[:, :, 1], y[:, :, 0] - y[:, :, 1]], dim=2).reshape(-1, d)
        h *= 2
    return (y * d ** -0.5).view(x.shape).to(x.dtype)


def compute_index_scores(q: torch.Tensor, weights: torch.Tensor, k: torch.Tensor, use_relu: bool = True) -> torch.Tensor:
    """q [sq, b, h, d], weights [sq, b, h], k [sk, b, d] → fp32 [b, sq, sk]."""
    s = torch.einsum("sbhd,tbd->bsht", q.float(), k.float())
    if use_relu:
        s = torch.relu(s)
    return (s * weights.float().transpose(0, 1).unsqueeze(-1)).sum(dim=2)


def topk_causal_indices(index_scores: torch.Tensor, topk: int, q_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """→ (indices [b, sq, k] into the key axis, valid [b, sq, k] bool).  Query t may only select keys ≤ t + q_offset; rows with fewer than k causal
    keys get the surplus slots marked invalid."""
    b, sq, sk = index_scores.shape
    k = min(topk, sk)
    qpos = torch.arange(sq, device=index_scores.device)[:, None] + q_offset
    future = torch.arange(sk, device=index_scores.device)[None, :] > qpos
    masked = index_scores.masked_fill(future[None], float("-inf"))
    vals, idx = torch.topk(masked, k, dim=-1)
    return idx, torch.isfinite(vals)


def sparse_attention_topk(q_abs: torch.Tensor, kv: torch.Tensor, idx: torch.Tensor, valid: torch.Tensor, scale: float, v_width: int) -> torch.Tensor:
    """Attention of every query over ITS selected keys.  q_abs [sq, b, n, c], kv [sk, b, c], idx / valid [b, sq, k] → [sq, b, n, v_width]
    (values are the first ``v_width`` channels of the selected kv rows: the latent part)."""
    sq, b, n, c = q_abs.shape
    kk = idx.shape[-1]
    rows = kv.transpose(0, 1)                                                        # [b, sk, c]
    sel = torch.gather(rows.unsqueeze(1).expand(b, sq, rows.shape[1], c), 2, idx.unsqueeze(-1).expand(b, sq, kk, c))   # [b, sq, k, c]
    scores = torch.einsum("sbnc,bskc->bsnk", q_abs.float(), sel.float()) * scale
    scores = scores.masked_fill(~valid[:, :, None, :], float("-inf"))
    probs = torch.softmax(scores, dim=-1)
    out = torch.einsum("bsnk,bskv->sbnv", probs, sel[..., :v_width].float())
    return out.to(q_abs.dtype)


def compute_dsa_indexer_loss(index_scores: torch.Tensor, q_abs: torch.Tensor, kv: torch.Tensor, scale: float, loss_coeff: float,
                             idx: Optional[torch.Tensor] = None, valid: Optional[torch.Tensor] = None, q_offset: int = 0,
                             cu_seqlens: Optional[torch.Tensor] = None, query_valid_rows: Optional[torch.Tensor] = None,
                             calculate_per_token_loss: bool = False) -> torch.Tensor:
    """KL( p_attn ‖ softmax(I) ), where p_attn is the main branch's attention distribution summed over heads and renormalised (detached).
    With ``idx`` (sparse variant) both distributions are restricted to the selected keys; otherwise they cover all causal keys.
    ``cu_seqlens`` (packed batch, b = 1): visibility is additionally confined to the query's own sequence; ``query_valid_rows``
    [b, sq] removes padding rows from the average; ``calculate_per_token_loss`` returns the sum (the trainer normalises)."""
    from .dsa_indexer_loss import indexer_loss_from_target, normalize_indexer_target
    from .dsa_masking import build_valid_mask_from_starts_ends, generate_varlen_mask_params_for_positions
    b, sq, sk = index_scores.shape
    with torch.no_grad():
        att = torch.einsum("sbnc,tbc->bnst", q_abs.float(), kv.float()) * scale
        qpos = torch.arange(sq, device=att.device)[:, None] + q_offset
        future = torch.arange(sk, device=att.device)[None, :] > qpos
        if cu_seqlens is not None:
            st, en = generate_varlen_mask_params_for_positions(cu_seqlens, qpos.view(-1))
            future = ~build_valid_mask_from_starts_ends(st, en, torch.arange(sk, device=att.device))
        att = torch.softmax(att.masked_fill(future[None, None], float("-inf")), dim=-1).sum(dim=1)          # [b, sq, sk]
    logits = index_scores.masked_fill(future[None], float("-inf"))
    if idx is not None:
        att = torch.gather(att, 2, idx) * valid
        logits = torch.gather(logits, 2, idx).masked_fill(~valid, float("-inf"))
    visible = torch.isfinite(logits)
    logp = torch.log_softmax(logits, dim=-1)
    return indexer_loss_from_target(normalize_indexer_target(att), logp, loss_coeff, query_valid_rows, calculate_per_token_loss, visible)


class DSAIndexerLossAutoScaler(torch.autograd.Function):
    """Attach the indexer loss to an activation: identity in forward, contributes ``scale · ∂loss`` in backward (same device as the MoE aux loss)."""

    main_loss_backward_scale: torch.Tensor = torch.tensor(1.0)

    @staticmethod
    def forward(ctx, output, loss):
        ctx.save_for_backward(loss)
        return output

    @staticmethod
    def backward(ctx, g):
        (loss,) = ctx.saved_tensors
        return g, torch.ones_like(loss) * DSAIndexerLossAutoScaler.main_loss_backward_scale.to(loss.device)

    @staticmethod
    def set_loss_scale(scale):
        DSAIndexerLossAutoScaler.main_loss_backward_scale = scale if isinstance(scale, torch.Tensor) else torch.tensor(float(scale))


@dataclass
class DSAIndexerSubmodules:
    linear_wq_b: Union[ModuleSpec, type] = None
    linear_wk: Union[ModuleSpec, type] = None
    k_norm: Union[ModuleSpec, type] = None
    linear_weights_proj: Union[ModuleSpec, type] = None


@dataclass
class DSAttentionSubmodules:
    indexer: Union[ModuleSpec, type] = None


class DSAIndexer(MegatronModule):
    """Index branch: ``q_idx`` from the query latent (or the hidden state), ``k_idx`` and the per-head weights from the hidden state; the rope
    part of both is rotated with the layer's angles, the rest goes through the Hadamard rotation."""

    def __init__(self, config, submodules: Optional[DSAIndexerSubmodules] = None, q_in_features: Optional[int] = None, rope_dim: int = 0):
        super().__init__(config)
        c = config
        self.n_heads = getattr(c, "dsa_indexer_n_heads", None) or 4
        self.head_dim = getattr(c, "dsa_indexer_head_dim", None) or 32
        self.topk = getattr(c, "dsa_indexer_topk", None) or 64
        self.use_relu = getattr(c, "dsa_indexer_use_relu", True) and getattr(c, "dsa_indexer_scoring_relu", True)
        self.rotate = getattr(c, "dsa_indexer_rotate_activation", True)
        self.rope_interleaved = getattr(c, "dsa_indexer_rope_interleaved", False)
        self.k_norm_fp32 = getattr(c, "dsa_indexer_k_norm_fp32", False)
        self.rope_dim = rope_dim
        q_in = q_in_features or c.hidden_size
        dev = "cpu" if (c.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        lin = lambda i, o: torch.nn.Linear(i, o, bias=False, device=dev, dtype=c.params_dtype)  # noqa: E731  (replicated: the index branch is tiny)
        self.linear_wq_b = lin(q_in, self.n_heads * self.head_dim)
        self.linear_wk = lin(c.hidden_size, self.head_dim)
        self.k_norm = torch.nn.LayerNorm(self.head_dim, eps=getattr(c, "dsa_indexer_k_norm_epsilon", None) or c.layernorm_epsilon, device=dev, dtype=c.params_dtype)
        self.linear_weights_proj = lin(c.hidden_size, self.n_heads)
        for m in (self.linear_wq_b, self.linear_wk, self.linear_weights_proj):
            c.init_method(m.weight)
        self.softmax_scale = self.head_dim ** -0.5

    def _rope(self, t, angles):
        had = rotate_activation if self.rotate else (lambda x: x)
        if self.rope_dim == 0 or angles is None:
            return had(t)
        from .... import ops

        rot, rest = t[..., : self.rope_dim], t[..., self.rope_dim :]
        rot = ops.apply_rope(rot.contiguous(), angles, self.rope_interleaved, 1.0)
        return had(torch.cat([rot, rest], dim=-1))

    def forward(self, hidden_states, q_source=None, angles=None, q_offset: int = 0, k_cache: Optional[torch.Tensor] = None):
        """hidden_states [s, b, h]; q_source defaults to the hidden state.  → (index_scores [b, sq, sk], topk idx, valid, k_idx)."""
        s, b = hidden_states.shape[:2]
        q = self.linear_wq_b(hidden_states if q_source is None else q_source).view(s, b, self.n_heads, self.head_dim)
        k = self.linear_wk(hidden_states)
        k = (torch.nn.functional.layer_norm(k.float(), (self.head_dim,), self.k_norm.weight.float(), self.k_norm.bias.float(), self.k_norm.eps).to(k.dtype)
             if self.k_norm_fp32 else self.k_norm(k)).unsqueeze(2)
        q, k = self._rope(q, angles), self._rope(k, angles).squeeze(2)
        if k_cache is not None:
            k = torch.cat([k_cache, k], dim=0)
        w = self.linear_weights_proj(hidden_states).float() * (self.n_heads ** -0.5) * self.softmax_scale
        scores = compute_index_scores(q, w, k, self.use_relu)
        idx, valid = topk_causal_indices(scores, self.topk, q_offset)
        return scores, idx, valid, k


class DSAttention(MegatronModule):
    """Sparse core attention on the absorbed latent.  ``forward(q_abs, kv, hidden_states, ...)`` → [sq, b, n, r] (the caller applies W_UV)."""

    def __init__(self, config, submodules: Optional[DSAttentionSubmodules] = None, layer_number: int = 1, softmax_scale: Optional[float] = None,
                 q_in_features: Optional[int] = None, **kwargs):
        super().__init__(config)
        self.layer_number = layer_number
        self.softmax_scale = softmax_scale
        self.loss_coeff = getattr(config, "dsa_indexer_loss_coeff", 0.0)
        self.sparse_loss = getattr(config, "dsa_indexer_use_sparse_loss", False)
        # reference names first (``dsa_indexer_skip_topk_offset`` / ``dsa_indexer_topk_freq``), this framework's earlier short names as a fallback
        self.skip_offset = getattr(config, "dsa_indexer_skip_topk_offset", None) or getattr(config, "dsa_skip_topk_offset", 0) or 0
        self.topk_freq = getattr(config, "dsa_indexer_topk_freq", None) or getattr(config, "dsa_topk_freq", 1) or 1
        self.computes_topk = not is_dsa_skip_topk_layer(layer_number, self.skip_offset, self.topk_freq)
        ind = submodules.indexer if submodules is not None and submodules.indexer is not None else DSAIndexer
        self.indexer = build_module(ind, config, q_in_features=q_in_features) if self.computes_topk else None
        self.last_indices = None

    def forward(self, q_abs, kv, hidden_states, v_width: int, q_source=None, angles=None, q_offset: int = 0, shared_indices=None):
        if self.indexer is not None:
            # the indexer sees detached activations: its only training signal is the KL term, it must not steer the main branch
            scores, idx, valid, _ = self.indexer(hidden_states.detach(), None if q_source is None else q_source.detach(), angles, q_offset)
            self.last_indices = (idx, valid)
        else:
            assert shared_indices is not None, f"layer {self.layer_number} reuses top-k indices: pass the (idx, valid) of layer " \
                                               f"{source_dsa_compute_layer(self.layer_number, self.skip_offset, self.topk_freq)}"
            idx, valid = shared_indices
            scores = None
        out = sparse_attention_topk(q_abs, kv, idx, valid, self.softmax_scale, v_width)
        if self.training and scores is not None and self.loss_coeff > 0:
            loss = compute_dsa_indexer_loss(scores, q_abs.detach(), kv.detach(), self.softmax_scale, self.loss_coeff,
                                            idx if self.sparse_loss else None, valid if self.sparse_loss else None, q_offset)
            out = DSAIndexerLossAutoScaler.apply(out, loss)
        return out
