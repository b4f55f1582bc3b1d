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
