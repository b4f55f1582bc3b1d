"""Nested-elastic ("Flextron / Matryoshka") models — reference ``megatron/elastification/`` (3.8 kLoC): ONE set of weights that can be run at several sizes.

A budget fixes, per layer, how many FFN units, how many attention heads and which layers are active.  Sub-networks are NESTED: units are ranked once by
importance and a budget of ``k`` always keeps the ``k`` most important ones, so every smaller model is a prefix of every larger one and all of them can be
trained together (sandwich rule: largest + smallest + random budgets per step, the largest one teaching the others by distillation).

Mechanics here (no model surgery — works on any ``TransformerLayer``-based model of this framework, under TP because masks act on the LOCAL shard of the
unit axis):

* ``ElasticController.calibrate`` runs a few batches with hooks on ``mlp.linear_fc2`` / ``self_attention.linear_proj`` inputs and accumulates per-unit /
  per-head activation energy → importance ranks (registered as buffers, so they are checkpointed with the model);
* ``ElasticController.set_budget`` installs masks: a forward pre-hook multiplies the fc2 input by the unit mask and the proj input by the head mask, a pair
  of layer hooks turns a disabled layer into the identity;
* ``extract_mlp_subnetwork`` materialises a budget as physically smaller fc1 / fc2 weights (deployment), numerically identical to the masked model."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class ElasticBudget:
    ffn_fraction: float = 1.0          # share of FFN units kept in every layer
    head_fraction: float = 1.0         # share of attention heads kept
    layer_fraction: float = 1.0        # share of layers kept (least important dropped first)

    def key(self):
        return (round(self.ffn_fraction, 4), round(self.head_fraction, 4), round(self.layer_fraction, 4))


class BudgetSampler:
    """Sandwich rule over a discrete set of choices."""

    def __init__(self, ffn: Sequence[float] = (0.25, 0.5, 0.75, 1.0), heads: Sequence[float] = (0.5, 1.0), layers: Sequence[float] = (1.0,), n_random: int = 1, seed: int = 0):
        self.ffn, self.heads, self.layers, self.n_random = sorted(ffn), sorted(heads), sorted(layers), n_random
        self.rng = random.Random(seed)

    def largest(self) -> ElasticBudget:
        return ElasticBudget(self.ffn[-1], self.heads[-1], self.layers[-1])

    def smallest(self) -> ElasticBudget:
        return ElasticBudget(self.ffn[0], self.heads[0], self.layers[0])

    def sample(self) -> List[ElasticBudget]:
        out = [self.largest(), self.smallest()]
        for _ in range(self.n_random):
            out.append(ElasticBudget(self.rng.choice(self.ffn), self.rng.choice(self.heads), self.rng.choice(self.layers)))
        return out


class ElasticController:
    def __init__(self, model: torch.nn.Module):
        from ..core.transformer.transformer_layer import TransformerLayer

        self.model = model
        self.layers = [m for m in model.modules() if isinstance(m, TransformerLayer)]
        assert self.layers, "no TransformerLayer found"
        self.budget = ElasticBudget()
        self._handles: List = []
        for L in self.layers:
            fc2, proj = L.mlp.linear_fc2, L.self_attention.linear_proj
            ffn = fc2.weight.shape[1]
            nh = L.self_attention.num_attention_heads_per_partition
            dev = fc2.weight.device
            # rank[i] = position of unit i in the importance order (0 = most important); identity until calibrated
            L.register_buffer("elastic_ffn_rank", torch.arange(ffn, device=dev), persistent=True)
            L.register_buffer("elastic_head_rank", torch.arange(nh, device=dev), persistent=True)
        self.register_layer_rank(torch.arange(len(self.layers)))
        self._install()

    def register_layer_rank(self, rank: torch.Tensor) -> None:
        self.layer_rank = rank.clone()

    # ---- hooks ----
    def _install(self) -> None:
        for li, L in enumerate(self.layers):
            def fc2_pre(mod, args, _L=L):
                k = max(1, int(round(self.budget.ffn_fraction * _L.elastic_ffn_rank.numel())))
                if k >= _L.elastic_ffn_rank.numel():
                    return None
                mask = (_L.elastic_ffn_rank < k).to(args[0].dtype)
                return (args[0] * mask,) + tuple(args[1:])

            def proj_pre(mod, args, _L=L):
                nh = _L.elastic_head_rank.numel()
                k = max(1, int(round(self.budget.head_fraction * nh)))
                if k >= nh:
                    return None
                d = args[0].shape[-1] // nh
                mask = (_L.elastic_head_rank < k).to(args[0].dtype).repeat_interleave(d)
                return (args[0] * mask,) + tuple(args[1:])

            def layer_pre(mod, args, kwargs, _li=li):
                mod._elastic_in = args[0] if args else kwargs.get("hidden_states")
                return None

            def layer_post(mod, args, kwargs, out, _li=li):
                keep = max(1, int(round(self.budget.layer_fraction * len(self.layers))))
                if self.layer_rank[_li] < keep:
                    return None
                return (mod._elastic_in,) + tuple(out[1:]) if isinstance(out, tuple) else mod._elastic_in

            self._handles += [L.mlp.linear_fc2.register_forward_pre_hook(fc2_pre), L.self_attention.linear_proj.register_forward_pre_hook(proj_pre),
                              L.register_forward_pre_hook(layer_pre, with_kwargs=True), L.register_forward_hook(layer_post, with_kwargs=True)]

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    def set_budget(self, budget: ElasticBudget) -> None:
        self.budget = budget

    # ---- importance ----
    @torch.no_grad()
    def calibrate(self, batches, forward_fn) -> None:
        """``forward_fn(model, batch)`` runs one forward.  Activation energy per FFN unit / head, and per-layer residual contribution, define the nested order."""
        acc_f = [torch.zeros_like(L.elastic_ffn_rank, dtype=torch.float32) for L in self.layers]
        acc_h = [torch.zeros_like(L.elastic_head_rank, dtype=torch.float32) for L in self.layers]
        acc_l = torch.zeros(len(self.layers))
        hs = []
        for li, L in enumerate(self.layers):
            def fc2_obs(m, a, _i=li):
                acc_f[_i].add_(a[0].float().pow(2).reshape(-1, a[0].shape[-1]).sum(0))

            hs.append(L.mlp.linear_fc2.register_forward_pre_hook(fc2_obs))

            def proj_obs(m, a, _i=li, _L=L):
                nh = _L.elastic_head_rank.numel()
                x = a[0].float().reshape(-1, nh, a[0].shape[-1] // nh)
                acc_h[_i].add_(x.pow(2).sum(dim=(0, 2)))

            hs.append(L.self_attention.linear_proj.register_forward_pre_hook(proj_obs))

            def layer_obs(m, args, kwargs, out, _i=li):
                o = out[0] if isinstance(out, tuple) else out
                acc_l[_i] += float((o.float() - m._elastic_in.float()).pow(2).mean() / m._elastic_in.float().pow(2).mean().clamp(min=1e-12))

            hs.append(L.register_forward_hook(layer_obs, with_kwargs=True))
        saved, self.budget = self.budget, ElasticBudget()
        was = self.model.training
        self.model.eval()
        try:
            for b in batches:
                forward_fn(self.model, b)
        finally:
            self.model.train(was)
            self.budget = saved
            for h in hs:
                h.remove()
        for L, f, h in zip(self.layers, acc_f, acc_h):
            L.elastic_ffn_rank.copy_(torch.argsort(torch.argsort(f, descending=True)))
            L.elastic_head_rank.copy_(torch.argsort(torch.argsort(h, descending=True)))
        self.layer_rank = torch.argsort(torch.argsort(acc_l, descending=True))

    def active_parameter_fraction(self, budget: Optional[ElasticBudget] = None) -> float:
        b = budget or self.budget
        tot = act = 0
        keep_layers = max(1, int(round(b.layer_fraction * len(self.layers))))
        for li, L in enumerate(self.layers):
            mlp = sum(p.numel() for p in L.mlp.parameters())
            att = sum(p.numel() for p in L.self_attention.parameters())
            tot += mlp + att
            if self.layer_rank[li] < keep_layers:
                act += mlp * b.ffn_fraction + att * b.head_fraction
        return act / max(tot, 1)


def distillation_step(controller: ElasticController, sampler: BudgetSampler, batch, forward_logits, task_loss, temperature: float = 1.0, kd_weight: float = 1.0):
    """One sandwich-rule training step: the largest budget is trained on the task loss and (detached) teaches every other sampled budget through a KL term.
    ``forward_logits(model, batch) -> logits``; ``task_loss(logits, batch) -> scalar``.  Returns the summed loss (call ``.backward()`` on it)."""
    budgets = sampler.sample()
    controller.set_budget(budgets[0])
    big = forward_logits(controller.model, batch)
    loss = task_loss(big, batch)
    teacher = F.log_softmax(big.detach().float() / temperature, dim=-1)
    for b in budgets[1:]:
        if b.key() == budgets[0].key():
            continue
        controller.set_budget(b)
        small = forward_logits(controller.model, batch)
        kd = F.kl_div(F.log_softmax(small.float() / temperature, dim=-1), teacher, log_target=True, reduction="batchmean") * temperature * temperature
        loss = loss + kd_weight * kd + task_loss(small, batch)
    controller.set_budget(budgets[0])
    return loss


@torch.no_grad()
def extract_mlp_subnetwork(layer, ffn_fraction: float) -> Dict[str, torch.Tensor]:
    """Physically smaller fc1 / fc2 weights for one layer at the given budget (gate and up halves of a gated fc1 are sliced consistently)."""
    rank = layer.elastic_ffn_rank
    ffn = rank.numel()
    k = max(1, int(round(ffn_fraction * ffn)))
    idx = torch.nonzero(rank < k).flatten()
    fc1, fc2 = layer.mlp.linear_fc1.weight, layer.mlp.linear_fc2.weight
    gated = fc1.shape[0] == 2 * ffn
    rows = torch.cat([idx, idx + ffn]) if gated else idx
    out = {"linear_fc1.weight": fc1[rows].clone(), "linear_fc2.weight": fc2[:, idx].clone()}
    if getattr(layer.mlp.linear_fc1, "bias", None) is not None:
        out["linear_fc1.bias"] = layer.mlp.linear_fc1.bias[rows].clone()
    return out
