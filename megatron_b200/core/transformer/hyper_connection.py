"""Hyper-connections (Zhu et al. 2024) — learnable replacement of the residual connection (reference ``transformer/hyper_connection.py``).

The hidden state is widened to ``n`` parallel residual streams ``H ∈ R^{n×d}`` per token.  Around a layer ``f`` (attention or MLP):

    x        = Σ_i  α^{in}_i · H_i                      (width connection: mix the streams into the layer input)
    H'_j     = Σ_i  A_{ij} · H_i  +  β_j · f(x)         (depth connection: stream-to-stream mixing + write the layer output)

with ``α^{in} ∈ R^n, A ∈ R^{n×n}, β ∈ R^n`` = static parameters + (optional) dynamic, input-dependent corrections
``tanh(norm(H) W) · s``.  Initialisation reproduces the pre-norm residual network: stream ``k = layer_idx mod n`` feeds the layer, A = I,
β = 1.  ``expand_streams`` / ``reduce_streams`` convert between ``[s, b, d]`` and ``[s, b, n, d]`` at the model boundary."""
from __future__ import annotations

import torch

from .module import MegatronModule


def expand_streams(x: torch.Tensor, n: int) -> torch.Tensor:
    """[s, b, d] → [s, b, n, d] (every stream starts as a copy of the embedding)."""
    return x.unsqueeze(2).expand(-1, -1, n, -1).contiguous()


def reduce_streams(h: torch.Tensor) -> torch.Tensor:
    """[s, b, n, d] → [s, b, d] (sum of the streams feeds the final norm)."""
    return h.sum(dim=2)


class HyperConnection(MegatronModule):
    def __init__(self, config, num_streams: int, layer_index: int, dynamic: bool = True, hidden_size: int = None):
        super().__init__(config)
        n, d = num_streams, hidden_size or config.hidden_size
        self.n, self.dynamic = n, dynamic
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        k = layer_index % n
        alpha_in = torch.zeros(n)
        alpha_in[k] = 1.0
        self.static_alpha_in = torch.nn.Parameter(alpha_in.to(dev))                 # [n]
        self.static_A = torch.nn.Parameter(torch.eye(n, device=dev))                 # [n, n]
        self.static_beta = torch.nn.Parameter(torch.ones(n, device=dev))             # [n]
        if dynamic:
            self.dyn_alpha_proj = torch.nn.Parameter(torch.zeros(d, n + n * n, device=dev))
            self.dyn_beta_proj = torch.nn.Parameter(torch.zeros(d, n, device=dev))
            self.dyn_alpha_scale = torch.nn.Parameter(torch.full((1,), 0.01, device=dev))
            self.dyn_beta_scale = torch.nn.Parameter(torch.full((1,), 0.01, device=dev))
        for p in self.parameters():
            setattr(p, "sequence_parallel", bool(config.sequence_parallel))  # replicated, fed by sequence-sharded activations

    def _coeffs(self, h: torch.Tensor):
        n = self.n
        a_in, A, beta = self.static_alpha_in, self.static_A, self.static_beta
        if not self.dynamic:
            return a_in, A, beta
        hn = torch.nn.functional.rms_norm(h.float(), (h.shape[-1],))
        pooled = hn.mean(dim=2)                                                      # [s, b, d]
        da = torch.tanh(pooled @ self.dyn_alpha_proj.float()) * self.dyn_alpha_scale
        db = torch.tanh(pooled @ self.dyn_beta_proj.float()) * self.dyn_beta_scale
        a_in = a_in + da[..., :n]                                                    # [s, b, n]
        A = A + da[..., n:].view(*da.shape[:-1], n, n)                               # [s, b, n, n]
        beta = beta + db                                                              # [s, b, n]
        return a_in, A, beta

    def width_connection(self, h: torch.Tensor):
        """h [s, b, n, d] → (layer input [s, b, d], saved coefficients)."""
        a_in, A, beta = self._coeffs(h)
        x = torch.einsum("...n,...nd->...d", a_in.expand(*h.shape[:-1]).to(h.dtype) if a_in.dim() == 1 else a_in.to(h.dtype), h)
        return x, (A, beta)

    def depth_connection(self, h: torch.Tensor, layer_out: torch.Tensor, saved):
        """h [s, b, n, d], layer_out [s, b, d] → new streams [s, b, n, d]."""
        A, beta = saved
        if A.dim() == 2:
            mixed = torch.einsum("ij,sbid->sbjd", A.to(h.dtype), h)
            return mixed + beta.to(h.dtype).view(1, 1, -1, 1) * layer_out.unsqueeze(2)
        mixed = torch.einsum("sbij,sbid->sbjd", A.to(h.dtype), h)
        return mixed + beta.to(h.dtype).unsqueeze(-1) * layer_out.unsqueeze(2)

    def forward(self, h: torch.Tensor, layer_fn):
        x, saved = self.width_connection(h)
        return self.depth_connection(h, layer_fn(x), saved)


# ---- manifold-constrained hyper-connections (mHC) -----------------------------------------------------------------------------------------
# reference ``transformer/hyper_connection.py:54-712`` (DeepSeek's mHC): the residual stream is n-wide, ``x_{l+1} = H_res x_l + H_postᵀ F(H_pre x_l)``, and the
# stream-mixing matrix H_res is constrained to the Birkhoff polytope (doubly stochastic: it mixes but never amplifies), obtained from unconstrained logits by
# Sinkhorn–Knopp iterations.  All three mappings are input-dependent: one projection of the RMS-scaled n·C-wide state gives n + n + n² numbers per token.


def sinkhorn_normalize(m: torch.Tensor, iterations: int, eps: float = 1e-6) -> torch.Tensor:
    """Alternate row / column normalisation of a positive matrix [..., n, n]."""
    for _ in range(iterations):
        m = m / m.sum(dim=-1, keepdim=True).clamp(min=eps)
        m = m / m.sum(dim=-2, keepdim=True).clamp(min=eps)
    return m


class SinkhornKnopp(torch.autograd.Function):
    """logits [..., n, n] → doubly stochastic matrix.  Only ``exp(logits − rowmax)`` is kept for the backward, which re-runs the iterations under autograd
    (n is 2–8: the iterations are cheap, the activations of 20 iterations per token per layer are not).  The row-max shift does not change the result: the
    first row normalisation cancels any per-row factor."""

    @staticmethod
    def forward(ctx, logits, iterations: int):
        m0 = torch.exp(logits - logits.max(dim=-1, keepdim=True).values)
        ctx.save_for_backward(m0)
        ctx.iterations = iterations
        return sinkhorn_normalize(m0, iterations)

    @staticmethod
    def backward(ctx, g):
        (m0,) = ctx.saved_tensors
        with torch.enable_grad():
            leaf = m0.detach().requires_grad_(True)
            out = sinkhorn_normalize(leaf, ctx.iterations)
            (gm0,) = torch.autograd.grad(out, leaf, g)
        return gm0 * m0, None            # d exp(x) / dx = exp(x)


class HyperConnectionModule(MegatronModule):
    """One mHC site (there are two per transformer layer: around attention and around the MLP)."""

    def __init__(self, config, layer_number: int):
        super().__init__(config)
        self.layer_number = layer_number
        self.n = n = config.mhc_num_residual_streams
        self.hidden_size = c = config.hidden_size
        self.sinkhorn_iterations = config.mhc_sinkhorn_iterations
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.mapping_proj = torch.nn.Linear(n * c, n * n + 2 * n, bias=False, device=dev, dtype=config.params_dtype)
        a = config.mhc_init_gating_factor
        self.alpha_pre = torch.nn.Parameter(torch.full((1,), a, device=dev))
        self.alpha_post = torch.nn.Parameter(torch.full((1,), a, device=dev))
        self.alpha_res = torch.nn.Parameter(torch.full((1,), a, device=dev))
        self.bias = torch.nn.Parameter(torch.zeros(n * n + 2 * n, device=dev))
        self.norm_eps = 1e-6
        if getattr(config, "perform_initialization", True):
            torch.nn.init.xavier_uniform_(self.mapping_proj.weight)
        for p in self.parameters():
            setattr(p, "sequence_parallel", bool(config.sequence_parallel))     # replicated over TP, fed by sequence-sharded activations

    # ---- mappings ----------------------------------------------------------------------------------------------------------------------
    def compute_mappings(self, x: torch.Tensor):
        """x [s, b, n·C] → h_pre [s, b, n] ∈ (0, 1), h_post [s, b, n] ∈ (0, 2), h_res [s, b, n, n] doubly stochastic."""
        n = self.n
        s, b, nc = x.shape
        r = 1.0 / (x.float().norm(dim=-1, keepdim=True) / (nc ** 0.5) + self.norm_eps)
        proj = self.mapping_proj(x).float()
        alpha = torch.cat([self.alpha_pre.expand(n), self.alpha_post.expand(n), self.alpha_res.expand(n * n)]).float()
        h = r * proj * alpha + self.bias.float()
        h_pre = torch.sigmoid(h[..., :n])
        h_post = 2.0 * torch.sigmoid(h[..., n:2 * n])
        h_res = SinkhornKnopp.apply(h[..., 2 * n:].reshape(s, b, n, n), self.sinkhorn_iterations)
        return h_pre, h_post, h_res

    def aggregate(self, x: torch.Tensor, h_pre: torch.Tensor) -> torch.Tensor:
        """n streams → the layer input: Σ_i h_pre_i · x_i."""
        s, b, nc = x.shape
        return torch.einsum("sbn,sbnc->sbc", h_pre.to(x.dtype), x.view(s, b, self.n, nc // self.n))

    def apply_h_post(self, y: torch.Tensor, h_post: torch.Tensor) -> torch.Tensor:
        """layer output [s, b, C] → n streams [s, b, n·C]: stream i receives h_post_i · y."""
        s, b, c = y.shape
        return (h_post.to(y.dtype).unsqueeze(-1) * y.unsqueeze(2)).reshape(s, b, self.n * c)

    def apply_h_res(self, h_res: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """stream mixing: out_i = Σ_j H_res[i, j] · x_j."""
        s, b, nc = residual.shape
        return torch.einsum("sbij,sbjc->sbic", h_res.to(residual.dtype), residual.view(s, b, self.n, nc // self.n)).reshape(s, b, nc)

    def forward(self, hidden_states: torch.Tensor):
        """→ (layer input [s, b, C], h_res, h_post)."""
        h_pre, h_post, h_res = self.compute_mappings(hidden_states)
        return self.aggregate(hidden_states, h_pre), h_res, h_post

    def fused_h_res_h_post_bda(self, h_res, residual, h_post, x_with_bias, dropout_prob: float, training: bool) -> torch.Tensor:
        """``H_res · residual + H_postᵀ · dropout(x + bias)`` — the bias-dropout-add of an mHC layer."""
        x, bias = x_with_bias
        if bias is not None:
            x = x + bias
        if dropout_prob > 0.0 and training:
            x = torch.nn.functional.dropout(x, p=dropout_prob, training=True)
        return self.apply_h_res(h_res, residual) + self.apply_h_post(x, h_post)

    @staticmethod
    def input_expand(x: torch.Tensor, n: int) -> torch.Tensor:
        """[s, b, C] → [s, b, n·C]: every stream starts as a copy of the embedding."""
        s, b, c = x.shape
        return x.unsqueeze(2).expand(s, b, n, c).reshape(s, b, n * c)

    @staticmethod
    def output_contract(x: torch.Tensor, n: int) -> torch.Tensor:
        """[s, b, n·C] → [s, b, C]: mean over the streams (so that contract(expand(x)) == x)."""
        s, b, nc = x.shape
        return x.view(s, b, n, nc // n).mean(dim=2)
